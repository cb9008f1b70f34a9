"""GPU: this library against the reference's OWN CUDA solvers.

parallel-implementation/L-BFGS{,-Wolfe,-Interpolation,-Backtracking,-Backtracking_Wolfe}.cu are cross-compiled
unmodified for sm_100 by `make -C oracle cudaref` (oracle/_ref/libref_cuda_*.so; the reference's own command line
except -arch and the renamed main()).  tests/golden/cuda_reference_traces.json holds their outputs from a B200
(oracle/make_golden_cuda.py); where the .so files are present the reference solver is also run live, on the same
GPU and in the same process as this library.

profile=CUDA selects the CUDA tree's outer loop; flavor=PAR is par/line_search.cpp (what par/L-BFGS.cu calls),
flavor=PAR_INLINED the copies inlined in the other four solvers.  The bar is BASELINE.json's: iterates within
1e-10 over 20 iterations, identical line-search decisions.
"""
import numpy as np
import pytest

from conftest import unhex
from test_oracle import cuda_case_is_pinned, cuda_case_setup

pytestmark = pytest.mark.gpu


def _sample(x, n):
    return x[:: max(1, n // 64)][:64]


def _trials_of_reference(case, K, want):
    """Distinct trial points the reference evaluated, from its count of f() calls: every iteration also calls
    f once before the search ("f(x_k)") and once for the "Optimum value" print; the solvers with an inlined
    Wolfe/interpolation search call f(x0) once more up front (par/L-BFGS-Wolfe.cu:172)."""
    if case["variant"] == "host" and case["line_search"] == "backtracking":
        # par/line_search.cpp:31 re-evaluates f(x) in every loop test: two calls per trial, plus the print
        return (want["f_evals"] - K) // 2
    extra = 0 if case["variant"] in ("host", "backtracking") else 1
    return want["f_evals"] - extra - 2 * K


@pytest.mark.parametrize("direction,graph", [("two_loop", 0), ("two_loop", 1), ("compact", 0)])
def test_solver_reproduces_the_cuda_reference_goldens(gpu, cuda_golden, direction, graph):
    worst = 0.0
    for name, case in cuda_golden["traces"].items():
        ls, flavor = cuda_case_setup(case)
        n = case["n"]
        x0 = gpu.x0_uniform(n, case["lo"], case["hi"])
        assert x0[0] == unhex(case["x0_first"]) and x0[-1] == unhex(case["x0_last"])
        for K, want in case["steps"].items():
            K = int(K)
            if not cuda_case_is_pinned(name, K):
                continue
            x, info, _ = gpu.solve(case["objective"], x0, ls, flavor, profile="cuda", m=case["m"], max_iterations=K,
                                   tolerance=case["tolerance"], direction=direction, use_graph=graph)
            err = float(np.max(np.abs(_sample(x, n) - np.array([unhex(v) for v in want["x_sample"]]))))
            worst = max(worst, err)
            assert err <= 1e-10, (name, K, direction, err)
            f_ref = unhex(want["f"])
            assert abs(info["f"] - f_ref) <= 1e-8 * max(1.0, abs(f_ref)), (name, K)
            if name != "host_interpolation_quad_1e4" and info["iterations"] == K:
                assert info["trial_evals"] == _trials_of_reference(case, K, want), (name, K, direction)
    assert worst < 1e-11, worst


@pytest.mark.parametrize("direction,graph", [("auto", 1), ("two_loop", 0)])
def test_run_to_convergence_matches_the_cuda_reference(gpu, cuda_golden, direction, graph):
    """profile=CUDA run to the tolerance (the `norm_g <= tolerance` exit, par/L-BFGS.cu:353-357) against the finals the
    reference's own CUDA programs reached on a B200 (oracle/make_golden_cuda.py, starts near the minimiser where the
    iteration count is stable): BASELINE.json's bar -- same verdict, iteration count within +-1, final f and |g| within
    1e-8 relative (values that are cancellation residues of f* = 0 get the summation noise of their terms as floor)."""
    assert cuda_golden.get("finals"), "tests/golden/cuda_reference_traces.json carries no finals: regenerate it (scripts/gpu_r2.sh cudagolden)"
    for name, case in cuda_golden["finals"].items():
        ls, flavor = cuda_case_setup(case)
        n = case["n"]
        x0 = gpu.x0_uniform(n, case["lo"], case["hi"])
        x, info, _ = gpu.solve(case["objective"], x0, ls, flavor, profile="cuda", m=case["m"], max_iterations=case["max_iterations"],
                               tolerance=case["tolerance"], direction=direction, use_graph=graph)
        tag = (name, direction, info["iterations"], case["iterations"])
        assert info["status"] == case["status"] == 0, tag
        assert abs(info["iterations"] - case["iterations"]) <= 1, tag
        f_ref, g_ref = unhex(case["f"]), unhex(case["gnorm"])
        if info["iterations"] == case["iterations"]:
            eps_f = n * 2.0 ** -52 * max(1.0, info["f0"]) * 1e-6  # f is a sum of n terms that started at f0 / n each
            assert abs(info["f"] - f_ref) <= 1e-8 * abs(f_ref) + eps_f, tag + (info["f"], f_ref)
            assert abs(info["gnorm"] - g_ref) <= 1e-8 * g_ref + 1e-12 * info["gnorm0"], tag + (info["gnorm"], g_ref)
            assert abs(x[0] - unhex(case["x_first"])) <= 1e-8 and abs(x[n // 2] - unhex(case["x_mid"])) <= 1e-8, tag
        assert info["gnorm"] <= case["tolerance"], tag


def test_live_cuda_reference_side_by_side(gpu, oracle, cuda_golden):
    """The reference's CUDA solver and this library on the same device, same process, on inputs that are NOT in
    the golden file; plus one golden case re-run live (the .so is the program that produced the file)."""
    from oracle import CudaRef
    if not CudaRef.available("wolfe"):
        pytest.skip("oracle/_ref/libref_cuda_*.so not built (make -C oracle cudaref in the build container)")
    # (variant, line-search argument, objective, n, (lo, hi), seed, m, K)
    cases = [("wolfe", None, "rosenbrock", 20011, (-2, 2), 7, 10, 20),
             ("wolfe", None, "rosenbrock", 3000, (-3, 3), 11, 4, 40),
             ("interpolation", None, "rosenbrock", 9973, (-2, 2), 5, 7, 30),
             ("backtracking", None, "rosenbrock", 12000, (-2, 2), 3, 10, 25),
             ("btwolfe", None, "rosenbrock", 5000, (-2, 2), 9, 10, 20),
             ("host", "backtracking", "rosenbrock", 8000, (-2, 2), 13, 10, 20),
             ("wolfe", None, "quadratic", 4096, (-1000, 1000), 2, 10, 3)]
    for variant, lsarg, obj, n, (lo, hi), seed, m, K in cases:
        case = dict(variant=variant, line_search=lsarg)
        ls, flavor = cuda_case_setup(case)
        x0 = gpu.x0_uniform(n, lo, hi, seed=seed)
        xr, ir = CudaRef(variant).lbfgs(obj, x0, lsarg or "wolfe", m, K, 0.0)
        x, info, tr = gpu.solve(obj, x0, ls, flavor, profile="cuda", m=m, max_iterations=K, tolerance=0.0, trace_rows=K)
        tag = (variant, lsarg, obj, n, K)
        assert np.max(np.abs(x - xr)) <= 1e-10, tag + (float(np.max(np.abs(x - xr))),)
        # every step the reference printed ("alpha: ...", 6 significant digits)
        k = info["iterations"]
        assert len(ir["alphas"]) >= k
        assert np.allclose(tr[:k, 3], ir["alphas"][:k], rtol=2e-6, atol=0), tag
        if obj != "quadratic":
            assert info["trial_evals"] == _trials_of_reference(case, K, ir), tag
        # and the CPU restatement agrees with both
        xo, io, _ = oracle.lbfgs(obj, x0, ls, flavor, m, K, 0.0, profile="cuda")
        assert np.max(np.abs(xo - xr)) <= 1e-10, tag
    # the golden file is what this program prints today
    case = cuda_golden["traces"]["wolfe_rosen_4097_m5"]
    x0 = gpu.x0_uniform(case["n"], case["lo"], case["hi"])
    xr, ir = CudaRef("wolfe").lbfgs("rosenbrock", x0, "wolfe", case["m"], 40, 0.0)
    want = case["steps"]["40"]
    assert np.max(np.abs(_sample(xr, case["n"]) - np.array([unhex(v) for v in want["x_sample"]]))) <= 1e-12
    assert (ir["f_evals"], ir["g_evals"]) == (want["f_evals"], want["g_evals"])
