"""Generate tests/golden/reference_traces.json from the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/make_golden.py

Every number comes from oracle/_ref/libref_{seq,hybrid}.so, i.e. the reference's own
sequential-implementation/lbfgs.cpp (+ its own or the CUDA tree's line_search.cpp) compiled
as-is by oracle/Makefile.  The reference returns only the final x, so the K-step state is
obtained by running it with max_iterations=K.  Doubles are stored as C99 hex floats (exact).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from oracle import Ref, build  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden", "reference_traces.json")

# (name, objective, n, (lo, hi), line_search, flavor, m, tol, [K...])
TRACE_CASES = [
    # BASELINE config 1 and its siblings (SURVEY.md 8(c) anchors)
    ("rosen_1e4_backtracking_seq", "rosenbrock", 10000, (-2, 2), "backtracking", "seq", 10, 1e-5, [1, 2, 5, 10, 20]),
    ("rosen_1e4_interpolation_seq", "rosenbrock", 10000, (-2, 2), "interpolation", "seq", 10, 1e-5, [1, 2, 5, 10, 20]),
    ("rosen_1e4_wolfe_seq", "rosenbrock", 10000, (-2, 2), "wolfe", "seq", 10, 1e-5, [1, 2, 5]),
    ("rosen_1e4_btwolfe_seq", "rosenbrock", 10000, (-2, 2), "backtracking_wolfe", "seq", 10, 1e-5, [1, 5, 20]),
    # hybrid oracle = seq outer loop + par/line_search.cpp (BASELINE config 2's semantics)
    ("rosen_1e4_wolfe_par", "rosenbrock", 10000, (-2, 2), "wolfe", "par", 10, 1e-5, [1, 2, 5, 10, 20, 50]),
    ("rosen_1e4_backtracking_par", "rosenbrock", 10000, (-2, 2), "backtracking", "par", 10, 1e-5, [1, 5, 20]),
    ("rosen_1e4_interpolation_par", "rosenbrock", 10000, (-2, 2), "interpolation", "par", 10, 1e-5, [1, 5, 20]),
    ("rosen_1e4_btwolfe_par", "rosenbrock", 10000, (-2, 2), "backtracking_wolfe", "par", 10, 1e-5, [1, 5, 20]),
    # history sweep of BASELINE config 3 (interpolation, m = 5/10/20), odd n on purpose
    ("rosen_4097_interp_m5", "rosenbrock", 4097, (-2, 2), "interpolation", "par", 5, 1e-5, [5, 20, 40]),
    ("rosen_4097_interp_m20", "rosenbrock", 4097, (-2, 2), "interpolation", "par", 20, 1e-5, [5, 20, 40]),
    ("quad_1e4_interp_m5", "quadratic", 10000, (-1000, 1000), "interpolation", "par", 5, 1e-8, [1, 2, 3]),
    ("tridiag_1e4_wolfe_par", "tridiag", 10000, (-2, 2), "wolfe", "par", 10, 1e-5, [1, 3, 6, 9]),
    ("tridiag_1e4_backtracking_seq", "tridiag", 10000, (-2, 2), "backtracking", "seq", 10, 1e-5, [1, 3, 6, 9]),
    # tiny and ragged sizes
    ("rosen_5_backtracking_seq", "rosenbrock", 5, (-2, 2), "backtracking", "seq", 10, 1e-5, [1, 5, 20]),
    ("rosen_2_wolfe_par", "rosenbrock", 2, (-2, 2), "wolfe", "par", 10, 1e-5, [1, 5, 20]),
    ("rosen_3_interp_seq", "rosenbrock", 3, (-2, 2), "interpolation", "seq", 3, 1e-5, [1, 5, 20]),
]

# run-to-convergence cases whose iteration count is stable
FINAL_CASES = [
    ("quad_1e4_backtracking_final", "quadratic", 10000, (-1000, 1000), "backtracking", "seq", 10, 1e-8, 15000),
    ("tridiag_1e4_backtracking_final", "tridiag", 10000, (-2, 2), "backtracking", "seq", 10, 1e-5, 1000),
    ("tridiag_1e4_interpolation_final", "tridiag", 10000, (-2, 2), "interpolation", "seq", 10, 1e-5, 1000),
    ("tridiag_1e4_wolfe_par_final", "tridiag", 10000, (-2, 2), "wolfe", "par", 10, 1e-5, 1000),
    ("rosen_5_backtracking_final", "rosenbrock", 5, (-2, 2), "backtracking", "seq", 10, 1e-5, 1000),
    ("rosen_1e4_near_backtracking_final", "rosenbrock", 10000, (0.5, 1.5), "backtracking", "seq", 10, 1e-5, 1000),
    ("rosen_1e4_near_wolfe_par_final", "rosenbrock", 10000, (0.5, 1.5), "wolfe", "par", 10, 1e-5, 1000),
]


def hx(v):
    return float(v).hex()


def main():
    build(ref=True)
    refs = {"seq": Ref("seq"), "par": Ref("par")}
    out = {"generator": "oracle/make_golden.py", "source": "oracle/_ref/libref_{seq,hybrid}.so "
           "(unmodified /root/reference sources, g++ -std=gnu++11 -O2 -ffp-contract=off, x86-64, no FMA)",
           "x0": "std::mt19937(42) + std::uniform_real_distribution<>(lo,hi)", "traces": {}, "finals": {}}
    for name, obj, n, (lo, hi), ls, flavor, m, tol, Ks in TRACE_CASES:
        ref = refs[flavor]
        x0 = ref.x0(n, lo, hi)
        steps = {}
        for K in Ks:
            x, info = ref.lbfgs(obj, x0, ls, m, K, tol)
            g = ref.grad(obj, x)
            steps[str(K)] = dict(f=hx(ref.f(obj, x)), gnorm=hx(ref.norm(g)), x_first=hx(x[0]),
                                 x_mid=hx(x[n // 2]), x_last=hx(x[-1]), x_sum=hx(float(np.sum(x))),
                                 status=info["status"], f_evals=info["f_evals"], g_evals=info["g_evals"])
        out["traces"][name] = dict(objective=obj, n=n, lo=lo, hi=hi, line_search=ls, flavor=flavor, m=m,
                                   tolerance=tol, x0_first=hx(x0[0]), x0_last=hx(x0[-1]), steps=steps)
        print(name, "ok")
    for name, obj, n, (lo, hi), ls, flavor, m, tol, max_it in FINAL_CASES:
        ref = refs[flavor]
        x0 = ref.x0(n, lo, hi)
        x, info = ref.lbfgs(obj, x0, ls, m, max_it, tol)
        g = ref.grad(obj, x)
        out["finals"][name] = dict(objective=obj, n=n, lo=lo, hi=hi, line_search=ls, flavor=flavor, m=m,
                                   tolerance=tol, max_iterations=max_it, status=info["status"],
                                   f=hx(ref.f(obj, x)), gnorm=hx(ref.norm(g)), f_evals=info["f_evals"],
                                   g_evals=info["g_evals"], x_first=hx(x[0]), x_mid=hx(x[n // 2]))
        print(name, "status", info["status"], "g_evals", info["g_evals"], "f", ref.f(obj, x))
    with open(OUT, "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    print("wrote", os.path.normpath(OUT))
    # the reference's own verbose stdout (seq/lbfgs.cpp:76-78 + the status line) for two small runs
    import ctypes as C
    ref = refs["seq"]
    ref.L.ref_lbfgs_verbose.restype = C.c_int
    for name, obj, ls, n, (lo, hi), max_it, m, tol in (("rosen5_backtracking", 1, 0, 5, (-2, 2), 12, 10, 1e-5),
                                                      ("tridiag64_interpolation", 2, 1, 64, (-2, 2), 20, 10, 1e-5)):
        x0 = ref.x0(n, lo, hi)
        buf = C.create_string_buffer(1 << 16)
        rc = ref.L.ref_lbfgs_verbose(obj, ls, C.c_size_t(n), x0.ctypes.data_as(C.POINTER(C.c_double)), max_it, m,
                                     C.c_double(tol), buf, C.c_size_t(1 << 16))
        assert rc > 0, rc
        path = os.path.join(HERE, "..", "tests", "golden", "verbose_%s.txt" % name)
        open(path, "w").write(buf.value.decode())
        print("wrote", os.path.normpath(path))


if __name__ == "__main__":
    main()
