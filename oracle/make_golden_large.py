"""Generate tests/golden/reference_large.json: parity goldens at the sizes BASELINE.json is quoted on.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/make_golden_large.py [--jobs J] [--only SUBSTR] [--skip-1e8] [--twins-only]

BASELINE config 3 (functions.cpp suite, n = 1e7, interpolation line search, m = 5/10/20) and config 2
(Rosenbrock n = 1e8, m = 10, Wolfe).  For every case and every checkpoint K:

  * the UNMODIFIED reference (oracle/_ref/libref_{seq,hybrid}.so) is run with max_iterations = K and its
    returned x is sampled at 256 evenly spaced indices; f(x), ||grad f(x)|| are evaluated with the
    reference's own functions; its f / gradient call counts are kept;
  * the C restatement (oracle/liblbfgs_oracle.so) is run once to the largest K with a per-iteration trace
    (f, ||g||, alpha, trials, history size, x[0], x[n/2]); its x at the largest K must be BIT-IDENTICAL to the
    reference's (recorded as `restatement_bitwise`; the generator aborts otherwise), which pins every trace
    row to the reference.

  * "exact_sums" twin: the same restatement with every reduction accumulated in long double (oracle_set_exact_sums, a
    TEST-ONLY switch).  It is the trajectory of the reference's algorithm WITHOUT the rounding noise of its naive
    left-to-right sums over 1e7 .. 1e8 terms.  The distance reference <-> twin (`spread`) is the reference's own
    summation noise; a solver with accurate (tree) sums lands on the twin, not on the reference, and is as far from
    the reference as the twin is.  tests/test_gpu_large.py asserts exactly that.

Doubles are stored as C99 hex floats.  The reference needs ~(2m+8) vectors of host memory: ~22 GB at n=1e8.
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from oracle import Oracle, Ref, build  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden", "reference_large.json")
NSAMPLE = 256

# (name, objective, n, (lo, hi), line_search, flavor, m, tolerance, [K...])
CASES = [
    # config 3: functions.cpp suite, n = 1e7, interpolation, history sweep (par/line_search.cpp constants)
    ("c3_rosen_1e7_interp_m5", "rosenbrock", 10_000_000, (-2, 2), "interpolation", "par", 5, 0.0, [5, 10, 20, 47]),
    ("c3_rosen_1e7_interp_m10", "rosenbrock", 10_000_000, (-2, 2), "interpolation", "par", 10, 0.0, [5, 10, 20]),
    ("c3_rosen_1e7_interp_m20", "rosenbrock", 10_000_000, (-2, 2), "interpolation", "par", 20, 0.0, [5, 10, 20]),
    ("c3_quad_1e7_interp_m5", "quadratic", 10_000_000, (-1000, 1000), "interpolation", "par", 5, 1e-5, [1, 2, 3]),
    ("c3_quad_1e7_interp_m10", "quadratic", 10_000_000, (-1000, 1000), "interpolation", "par", 10, 1e-5, [1, 2, 3]),
    ("c3_quad_1e7_interp_m20", "quadratic", 10_000_000, (-1000, 1000), "interpolation", "par", 20, 1e-5, [1, 2, 3]),
    # the sequential tree at the same size (config 1's line search, scaled up)
    ("seq_rosen_1e7_backtracking_m10", "rosenbrock", 10_000_000, (-2, 2), "backtracking", "seq", 10, 0.0, [5, 10, 20]),
    ("seq_rosen_1e7_interp_m10", "rosenbrock", 10_000_000, (-2, 2), "interpolation", "seq", 10, 0.0, [5, 10, 20]),
    # config 2: the headline workload
    ("c2_rosen_1e8_wolfe_m10", "rosenbrock", 100_000_000, (-2, 2), "wolfe", "par", 10, 0.0, [5, 10, 20]),
]


def hx(v):
    return float(v).hex()


def sample_index(n):
    return [(j * (n - 1)) // (NSAMPLE - 1) for j in range(NSAMPLE)]


def run_twin(case):
    """the exact-sums twin of a case: x samples, f, |g| at every checkpoint + the per-iteration trace"""
    name, obj, n, (lo, hi), ls, flavor, m, tol, Ks = case
    t_start = time.time()
    orc = Oracle()
    orc.set_exact_sums(True)
    x0 = orc.x0(n, lo, hi)
    idx = np.array(sample_index(n))
    steps = {}
    tr = None
    for K in Ks:
        x, info, tr = orc.lbfgs(obj, x0, ls, flavor, m, K, tol, trace_rows=K)
        g = orc.grad(obj, x)
        steps[str(K)] = dict(f=hx(orc.f(obj, x)), gnorm=hx(orc.norm(g)), status=info["status"], f_evals=info["f_evals"],
                             g_evals=info["g_evals"], x_absmax=hx(float(np.max(np.abs(x)))), x_sample=[hx(v) for v in x[idx]])
        del g
    trace = [dict(k=int(r[0]), f=hx(r[1]), gnorm=hx(r[2]), alpha=hx(r[3]), trials=int(r[4]), hist=int(r[5]),
                  x_first=hx(r[6]), x_mid=hx(r[7])) for r in tr]
    orc.set_exact_sums(False)
    print("%s exact-sums twin done in %.0f s" % (name, time.time() - t_start), flush=True)
    return name, dict(steps=steps, trace=trace,
                      what="oracle/liblbfgs_oracle.so with oracle_set_exact_sums(1): every reduction accumulated in long double")


def add_spread(rec):
    """distance reference <-> exact-sums twin at every checkpoint: the reference's own summation noise"""
    unhex = float.fromhex
    for K, tw in rec["exact_sums"]["steps"].items():
        ref = rec["steps"][K]
        a = np.array([unhex(v) for v in ref["x_sample"]])
        b = np.array([unhex(v) for v in tw["x_sample"]])
        fr, ft = unhex(ref["f"]), unhex(tw["f"])
        tw["spread_dx"] = float(np.max(np.abs(a - b)) / unhex(ref["x_absmax"]))
        tw["spread_df"] = abs(fr - ft) / abs(fr) if fr != 0 else abs(ft)


def run_case(case):
    name, obj, n, (lo, hi), ls, flavor, m, tol, Ks = case
    t_start = time.time()
    ref = Ref(flavor)
    x0 = ref.x0(n, lo, hi)
    idx = np.array(sample_index(n))
    steps = {}
    x_last = None
    for K in Ks:
        x, info = ref.lbfgs(obj, x0, ls, m, K, tol)
        g = ref.grad(obj, x)
        steps[str(K)] = dict(f=hx(ref.f(obj, x)), gnorm=hx(ref.norm(g)), status=info["status"], f_evals=info["f_evals"],
                             g_evals=info["g_evals"], seconds=info["seconds"], x_absmax=hx(float(np.max(np.abs(x)))),
                             x_sample=[hx(v) for v in x[idx]])
        del g
        x_last = x
        print("  %s K=%d  %.1f s" % (name, K, info["seconds"]), flush=True)
    del ref
    orc = Oracle()
    Kmax = max(Ks)
    xo, io, tr = orc.lbfgs(obj, x0, ls, flavor, m, Kmax, tol, trace_rows=Kmax)
    bitwise = bool(np.array_equal(xo, x_last))
    if not bitwise:
        raise SystemExit("%s: the restatement differs from the reference at K=%d (max |dx| %.3e)" %
                         (name, Kmax, float(np.max(np.abs(xo - x_last)))))
    trace = [dict(k=int(r[0]), f=hx(r[1]), gnorm=hx(r[2]), alpha=hx(r[3]), trials=int(r[4]), hist=int(r[5]),
                  x_first=hx(r[6]), x_mid=hx(r[7])) for r in tr]
    print("%s done in %.0f s (restatement bitwise: %s)" % (name, time.time() - t_start, bitwise), flush=True)
    return name, dict(objective=obj, n=n, lo=lo, hi=hi, line_search=ls, flavor=flavor, m=m, tolerance=tol,
                      x0_first=hx(x0[0]), x0_last=hx(x0[-1]), sample_index=[int(i) for i in idx], steps=steps,
                      trace=trace, restatement_bitwise=bitwise, restatement_iterations=int(io["iterations"]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs", type=int, default=4)
    ap.add_argument("--only", default="")
    ap.add_argument("--skip-1e8", action="store_true")
    ap.add_argument("--twins-only", action="store_true", help="keep the reference records of the existing file, (re)generate the exact-sums twins")
    args = ap.parse_args()
    build(ref=True)
    out = {"generator": "oracle/make_golden_large.py", "source": "oracle/_ref/libref_{seq,hybrid}.so (unmodified /root/reference "
           "sources, g++ -std=gnu++11 -O2 -ffp-contract=off, x86-64, no FMA); per-iteration trace rows from "
           "oracle/liblbfgs_oracle.so, bit-identical to the reference at the largest K of every case",
           "x0": "std::mt19937(42) + std::uniform_real_distribution<>(lo,hi)", "cases": {}}
    if os.path.exists(OUT):  # regenerate selectively
        out["cases"] = json.load(open(OUT)).get("cases", {})
    cases = [c for c in CASES if args.only in c[0] and not (args.skip_1e8 and c[2] > 20_000_000)]
    small = [c for c in cases if c[2] <= 20_000_000]
    big = [c for c in cases if c[2] > 20_000_000]
    if args.twins_only:
        small, big = [], []
    twins = [c for c in cases if c[1] != "quadratic"]  # the separable quadratic converges in one step: no noise to explain
    if small:
        with mp.Pool(min(args.jobs, len(small))) as pool:
            for name, rec in pool.imap_unordered(run_case, small):
                out["cases"][name] = rec
                json.dump(out, open(OUT, "w"), indent=0, sort_keys=True)
    for c in big:  # ~22 GB each: one at a time
        name, rec = run_case(c)
        out["cases"][name] = rec
        json.dump(out, open(OUT, "w"), indent=0, sort_keys=True)
    small_t = [c for c in twins if c[2] <= 20_000_000]
    big_t = [c for c in twins if c[2] > 20_000_000]
    if small_t:
        with mp.Pool(min(args.jobs, len(small_t))) as pool:
            for name, tw in pool.imap_unordered(run_twin, small_t):
                out["cases"][name]["exact_sums"] = tw
                add_spread(out["cases"][name])
                json.dump(out, open(OUT, "w"), indent=0, sort_keys=True)
    for c in big_t:
        name, tw = run_twin(c)
        out["cases"][name]["exact_sums"] = tw
        add_spread(out["cases"][name])
        json.dump(out, open(OUT, "w"), indent=0, sort_keys=True)
    print("wrote", os.path.normpath(OUT))


if __name__ == "__main__":
    main()
