"""Generate tests/golden/cuda_reference_traces.json from the reference's UNMODIFIED CUDA solvers.

TEST INFRASTRUCTURE ONLY.  Needs a GPU, so it runs on the GPU box (the .so files are cross-compiled in the
build container by `make -C oracle cudaref` and travel with the repo):

    gpurun -- 'python oracle/make_golden_cuda.py gpurun_out/cuda_reference_traces.json'
    cp gpurun_out/cuda_reference_traces.json tests/golden/

Every number comes from oracle/_ref/libref_cuda_*.so = parallel-implementation/L-BFGS*.cu + functions.cpp +
line_search.cpp + vector_utils.cpp compiled with the reference's own command line (par/run.sh) for sm_100,
executing on a B200 with the image's cuBLAS.  The reference returns only the final x, so the K-step state is
obtained by running it with max_iterations=K.  Doubles are stored as C99 hex floats (exact); f and ||g|| of the
returned x are evaluated with the reference's own host objective (oracle/_ref/libref_seq.so).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from oracle import CudaRef, Ref  # noqa: E402

# (name, variant, line_search argument (host variant only), objective, n, (lo, hi), m, tol, [K...])
CASES = [
    ("wolfe_rosen_1e4", "wolfe", None, "rosenbrock", 10000, (-2, 2), 10, 0.0, [1, 2, 3, 5, 10, 20]),
    ("wolfe_rosen_4097_m5", "wolfe", None, "rosenbrock", 4097, (-2, 2), 5, 0.0, [1, 5, 20, 40]),
    ("wolfe_quad_1e4", "wolfe", None, "quadratic", 10000, (-1000, 1000), 10, 0.0, [1, 2, 3]),
    ("backtracking_rosen_1e4", "backtracking", None, "rosenbrock", 10000, (-2, 2), 10, 0.0, [1, 2, 3, 5, 10, 20]),
    ("interpolation_rosen_1e4", "interpolation", None, "rosenbrock", 10000, (-2, 2), 10, 0.0, [1, 2, 3, 5, 10, 20]),
    ("interpolation_rosen_4097_m20", "interpolation", None, "rosenbrock", 4097, (-2, 2), 20, 0.0, [5, 20, 40]),
    ("btwolfe_rosen_1e4", "btwolfe", None, "rosenbrock", 10000, (-2, 2), 10, 0.0, [1, 2, 5, 10, 20]),
    ("host_wolfe_rosen_1e4", "host", "wolfe", "rosenbrock", 10000, (-2, 2), 10, 0.0, [1, 2, 3, 5, 10]),
    ("host_backtracking_rosen_1e4", "host", "backtracking", "rosenbrock", 10000, (-2, 2), 10, 0.0, [1, 2, 3, 5, 10]),
    ("host_interpolation_quad_1e4", "host", "interpolation", "quadratic", 10000, (-1000, 1000), 10, 0.0, [1, 2, 3]),
]

# run-to-convergence cases from starts near the minimiser, where the iteration count is stable under rounding
# (SURVEY.md App. D); the reference mains' own tolerance (1e-1) and a tighter one.
# (name, variant, line_search argument, objective, n, (lo, hi), m, tol, max_iterations)
FINAL_CASES = [
    ("wolfe_rosen_1e4_near_final", "wolfe", None, "rosenbrock", 10000, (0.5, 1.5), 10, 1e-1, 1000),
    ("wolfe_rosen_1e4_near_tight_final", "wolfe", None, "rosenbrock", 10000, (0.5, 1.5), 10, 1e-5, 1000),
    ("interpolation_rosen_1e4_near_final", "interpolation", None, "rosenbrock", 10000, (0.5, 1.5), 10, 1e-1, 1000),
    ("backtracking_rosen_1e4_near_final", "backtracking", None, "rosenbrock", 10000, (0.5, 1.5), 10, 1e-1, 1000),
    ("btwolfe_rosen_1e4_near_final", "btwolfe", None, "rosenbrock", 10000, (0.5, 1.5), 10, 1e-1, 1000),
    ("wolfe_quad_1e4_final", "wolfe", None, "quadratic", 10000, (-1000, 1000), 10, 1e-8, 100),
]


def hx(v):
    return float(v).hex()


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "..", "tests", "golden", "cuda_reference_traces.json")
    host = Ref("seq")
    out = {"generator": "oracle/make_golden_cuda.py",
           "source": "oracle/_ref/libref_cuda_*.so (unmodified /root/reference/parallel-implementation sources, "
                     "nvcc 12.9 defaults incl. -fmad=true, sm_100, cuBLAS of the image) executed on a B200",
           "x0": "std::mt19937(42) + std::uniform_real_distribution<>(lo,hi)", "traces": {}}
    for name, variant, ls, obj, n, (lo, hi), m, tol, Ks in CASES:
        ref = CudaRef(variant)
        x0 = host.x0(n, lo, hi)
        steps = {}
        log = ""
        for K in Ks:
            x, info = ref.lbfgs(obj, x0, ls or "wolfe", m, K, tol)
            g = host.grad(obj, x)
            steps[str(K)] = dict(f=hx(host.f(obj, x)), gnorm=hx(host.norm(g)), x_first=hx(x[0]), x_mid=hx(x[n // 2]),
                                 x_last=hx(x[-1]), x_sum=hx(float(np.sum(x))), status=info["status"],
                                 f_evals=info["f_evals"], g_evals=info["g_evals"],
                                 # 64 evenly spaced elements of the iterate, exact
                                 x_sample=[hx(v) for v in x[:: max(1, n // 64)][:64]])
            if K == max(Ks):
                steps[str(K)]["alphas_printed"] = info["alphas"]
                steps[str(K)]["gnorms_printed"] = info["gnorms"]
                log = "\n".join(l for l in info["log"].splitlines() if not l.startswith(("First x", "Found")))
        out["traces"][name] = dict(variant=variant, source_file=CudaRef.VARIANTS[variant], line_search=ls, objective=obj, n=n,
                                   lo=lo, hi=hi, m=m, tolerance=tol, x0_first=hx(x0[0]), x0_last=hx(x0[-1]), steps=steps,
                                   stdout_of_longest_run=log[:6000])
        print(name, "ok", "alphas", steps[str(max(Ks))]["alphas_printed"][:8], flush=True)
    out["finals"] = {}
    for name, variant, ls, obj, n, (lo, hi), m, tol, max_it in FINAL_CASES:
        ref = CudaRef(variant)
        x0 = host.x0(n, lo, hi)
        x, info = ref.lbfgs(obj, x0, ls or "wolfe", m, max_it, tol)
        g = host.grad(obj, x)
        out["finals"][name] = dict(variant=variant, source_file=CudaRef.VARIANTS[variant], line_search=ls, objective=obj, n=n,
                                   lo=lo, hi=hi, m=m, tolerance=tol, max_iterations=max_it, status=info["status"],
                                   iterations=len(info["gnorms"]), f=hx(host.f(obj, x)), gnorm=hx(host.norm(g)),
                                   f_evals=info["f_evals"], g_evals=info["g_evals"], x_first=hx(x[0]), x_mid=hx(x[n // 2]))
        print(name, "status", info["status"], "iterations", len(info["gnorms"]), flush=True)
    with open(out_path, "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    print("wrote", os.path.normpath(out_path))


if __name__ == "__main__":
    main()
