#!/usr/bin/env python
"""Measure the BASELINE.json configurations other than the headline bench line.

    python benchmarks/run_configs.py [config1] [config3] [msweep] [--gpus-note]

config1 : Rosenbrock n=1e4, m=10, backtracking (the reference's own CPU case) run to convergence,
          host-stepped vs CUDA-graph loop, next to the unmodified reference on one host core.
config3 : functions.cpp suite (quadratic, Rosenbrock) n=1e7, interpolation line search, m=5/10/20.
msweep  : history-size sweep at n=1e8 on one GPU, explicit two-loop vs compact form (config 5's
          sweep, single-GPU leg): iterations/s and fraction of the measured HBM peak per m.
Each result is printed as one JSON line.  (The CPU reference for config 1 is timed by
`bench.py --impl reference`, the only place besides tests/ that may execute oracle/.)
"""
import importlib.util
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_pkg():
    name = "cuda_lbfgs_b200"
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "cuda-lbfgs_b200", "__init__.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def steady(pkg, objective, n, x0, params, warm, steps, breakdown=None):
    s = pkg.Solver(objective, n, params, trace_rows=warm + steps + 2)
    s.set_x0(x0)
    s.iterate(warm)
    s.iterate(steps)
    r = s.result()
    tr = s.trace()
    done = int(r["iterations"]) - warm
    if breakdown is not None:
        _, classes = s.iterate_profiled(5)
        breakdown.update({k: round(v["ms"] / 5, 4) for k, v in classes.items() if v["launches"]})
    s.destroy()
    return r, tr, done


def config1(pkg):
    n = 10000
    x0 = pkg.x0_uniform(n, -2, 2)
    out = {"config": "1: Rosenbrock n=1e4, m=10, backtracking Armijo, tol 1e-5, run to convergence"}
    for graph, direction in ((0, "two_loop"), (1, "two_loop"), (0, "compact"), (1, "compact")):
        t = time.perf_counter()
        x, info, _ = pkg.solve("rosenbrock", x0, "backtracking", "seq", m=10, max_iterations=20000, tolerance=1e-5,
                               use_graph=graph, direction=direction)
        wall = time.perf_counter() - t
        out[direction + ("_graph" if graph else "_stepped")] = dict(iterations=info["iterations"], status=info["status"], f=info["f"],
                                                    gnorm=info["gnorm"], device_s=info["device_ms"] / 1e3, wall_s=wall,
                                                    iterations_per_s=info["iterations"] / (info["device_ms"] / 1e3),
                                                    launches=info["kernel_launches"])
    # the CPU side of this config is timed by `bench.py --impl reference` (key "config1"): only bench.py's
    # reference / cpu_baseline legs may execute oracle/
    print(json.dumps(out), flush=True)


def config3(pkg):
    n = 10_000_000
    pk = peak()
    for objective, (lo, hi) in (("quadratic", (-1000, 1000)), ("rosenbrock", (-2, 2))):
        x0 = pkg.x0_uniform(n, lo, hi)
        for m in (5, 10, 20):
            for direction in ("two_loop", "compact"):
                for graph in (0, 1):
                    p = pkg.default_params("par", line_search="interpolation", m=m, max_iterations=10 ** 9, tolerance=0.0,
                                           direction=direction, use_graph=graph)
                    warm, steps = (m + 2, 40) if objective == "rosenbrock" else (0, 3)
                    r, tr, done = steady(pkg, objective, n, x0, p, warm, steps)
                    if done <= 0:
                        continue
                    print(json.dumps({"config": "3: functions.cpp suite n=1e7, interpolation line search", "objective": objective,
                                      "m": m, "direction": direction, "graph": graph, "iterations": done,
                                      "iterations_per_s": done / (r["device_ms"] / 1e3), "ms_per_iteration": r["device_ms"] / done,
                                      "achieved_GBps": r["bytes_moved"] / (r["device_ms"] * 1e-3) / 1e9,
                                      "frac_of_measured_peak": r["bytes_moved"] / (r["device_ms"] * 1e-3) / 1e9 / pk,
                                      "f": r["f"], "status": r["status"]}), flush=True)


def msweep(pkg):
    n = 100_000_000
    pk = peak()
    x0 = pkg.x0_uniform(n, -2, 2)
    ms = [int(a) for a in os.environ.get("MSWEEP_M", "3,5,10,20,30,50").split(",")]
    dirs = os.environ.get("MSWEEP_DIR", "two_loop,compact").split(",")
    for m in ms:
        for direction in dirs:
            p = pkg.default_params("par", line_search="wolfe", m=m, max_iterations=10 ** 9, tolerance=0.0, direction=direction)
            bd = {}
            r, tr, done = steady(pkg, "rosenbrock", n, x0, p, m + 2, 15, bd)
            print(json.dumps({"config": "5 (1-GPU leg): history sweep n=1e8, Wolfe", "m": m, "direction": direction,
                              "class_ms_per_iteration": bd,
                              "iterations_per_s": done / (r["device_ms"] / 1e3), "ms_per_iteration": r["device_ms"] / done,
                              "algorithmic_GB_per_iteration": r["bytes_moved"] / done / 1e9,
                              "achieved_GBps": r["bytes_moved"] / (r["device_ms"] * 1e-3) / 1e9,
                              "frac_of_measured_peak": r["bytes_moved"] / (r["device_ms"] * 1e-3) / 1e9 / pk,
                              "frac_of_8TBps": r["bytes_moved"] / (r["device_ms"] * 1e-3) / 1e9 / 8000.0,
                              "trials_per_iteration": float(np.mean(tr[-done:, 4])), "f": r["f"]}), flush=True)


if __name__ == "__main__":
    pkg = load_pkg()
    which = [a for a in sys.argv[1:] if not a.startswith("-")] or ["config1", "config3", "msweep"]
    for w in which:
        {"config1": config1, "config3": config3, "msweep": msweep}[w](pkg)
