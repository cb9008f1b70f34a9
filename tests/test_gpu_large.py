"""GPU: parity at the sizes BASELINE.json is quoted on (config 3: n = 1e7, config 2: n = 1e8).

tests/golden/reference_large.json holds, for every case, what the UNMODIFIED reference returned after K = 5/10/20
iterations (f, ||g||, call counts and x at 256 evenly spaced indices) plus the per-iteration trace of the C
restatement, which is bit-identical to the reference at the largest K (oracle/make_golden_large.py).  Both direction
algorithms (explicit two-loop, compact/Gram) run in graph mode -- the configuration bench.py measures -- and must
meet BASELINE.json's bar: iterates within 1e-10 relative over the first 20 iterations, identical step decisions.
On a failure the first diverging iteration and the scalar that diverged are printed.  Every run appends what it
measured to gpurun_out/large_parity.jsonl (DESIGN.md quotes it).
"""
import json
import os

import numpy as np
import pytest

from conftest import ROOT, unhex

pytestmark = pytest.mark.gpu

TOL_ITERATE = 1e-10
GOLDEN = os.path.join(ROOT, "tests", "golden", "reference_large.json")


def _cases():
    if not os.path.exists(GOLDEN):
        return {}
    with open(GOLDEN) as fh:
        return json.load(fh)["cases"]


CASES = _cases()


def _first_divergence(tr, want_rows, upto):
    """(iteration, column name, got, want) of the first trace entry outside the bar, or None."""
    for k in range(min(upto, len(want_rows), len(tr))):
        w = want_rows[k]
        for col, key, tol in ((4, "trials", 0.0), (5, "hist", 0.0), (3, "alpha", 1e-9), (1, "f", TOL_ITERATE), (2, "gnorm", 1e-9)):
            ref = float(w[key]) if key in ("trials", "hist") else unhex(w[key])
            if abs(tr[k][col] - ref) > tol * abs(ref):
                return k, key, float(tr[k][col]), ref
    return None


@pytest.mark.parametrize("direction", ["two_loop", "compact"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_large_size_parity_with_the_reference(gpu, name, direction):
    case = CASES[name]
    n, m = case["n"], case["m"]
    Ks = sorted(int(k) for k in case["steps"])
    Kmax = Ks[-1]
    x0 = gpu.x0_uniform(n, case["lo"], case["hi"])
    assert x0[0] == unhex(case["x0_first"]) and x0[-1] == unhex(case["x0_last"])
    p = gpu.default_params(case["flavor"], line_search=case["line_search"], m=m, max_iterations=Kmax,
                           tolerance=case["tolerance"], direction=direction, use_graph=1)
    s = gpu.Solver(case["objective"], n, p, trace_rows=Kmax)
    s.set_x0(x0)
    del x0
    idx = np.asarray(case["sample_index"])
    report = {"case": name, "direction": direction, "n": n, "m": m, "checkpoints": {}}
    failures = []
    done = 0
    for K in Ks:
        want = case["steps"][str(K)]
        s.iterate(K - done)
        done = K
        r = s.result()
        xs = s.x()[idx]
        ref = np.array([unhex(v) for v in want["x_sample"]])
        dx = float(np.max(np.abs(xs - ref)) / unhex(want["x_absmax"]))
        df = abs(r["f"] - unhex(want["f"])) / abs(unhex(want["f"])) if unhex(want["f"]) != 0 else abs(r["f"])
        dg = abs(r["gnorm"] - unhex(want["gnorm"])) / max(abs(unhex(want["gnorm"])), 1e-300)
        report["checkpoints"][K] = {"max_rel_dx": dx, "rel_df": df, "rel_dgnorm": dg, "status": r["status"],
                                    "iterations": r["iterations"]}
        if want["status"] == 0:  # the reference converged before K: same verdict, iteration count within +-1
            if r["status"] != 0:
                failures.append((K, "status", r["status"], 0))
        elif K <= 20:
            if r["iterations"] != K:
                failures.append((K, "iterations", r["iterations"], K))
            if dx > TOL_ITERATE:
                failures.append((K, "max_rel_dx", dx, TOL_ITERATE))
            if df > TOL_ITERATE:
                failures.append((K, "rel_df", df, TOL_ITERATE))
    tr = s.trace()
    s.destroy()
    div = _first_divergence(tr, case["trace"], min(20, len(case["trace"])))
    div_any = _first_divergence(tr, case["trace"], len(case["trace"]))
    report["first_divergence_within_20"] = div
    report["first_divergence_any"] = div_any
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "large_parity.jsonl"), "a") as fh:
        fh.write(json.dumps(report) + "\n")
    print(json.dumps(report))
    assert div is None, "first diverging iteration %d: %s = %r, reference %r" % div
    assert not failures, failures
