#!/usr/bin/env python
"""Where the end-to-end time of one solve goes: create / set_x0 (H2D + graph build + f(x0)) / iterate / get_x (D2H) /
destroy, from pinned and from pageable host buffers, plus the one-shot lbfgsb200_solve().

    python benchmarks/e2e_breakdown.py [n] [iterations]

One JSON line per variant.  The first pass of a process pays one-off costs (pinned staging buffers, function attributes,
the memory pool's first mapping); passes 2 and 3 are the steady state.
"""
import importlib.util
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("cuda_lbfgs_b200", os.path.join(ROOT, "cuda-lbfgs_b200", "__init__.py"))
pkg = importlib.util.module_from_spec(spec)
sys.modules["cuda_lbfgs_b200"] = pkg
spec.loader.exec_module(pkg)
L = pkg.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
its = int(sys.argv[2]) if len(sys.argv) > 2 else 42
x0p = pkg.PinnedArray(n)
outp = pkg.PinnedArray(n)
pkg.x0_uniform(n, -2, 2, out=x0p.array)
x0g = np.array(x0p.array)   # pageable copies
outg = np.empty(n)
for kind, x0, out in (("pinned", x0p.array, outp.array), ("pageable", x0g, outg)):
    for rep in range(3):
        p = pkg.default_params("par", line_search="wolfe", m=10, max_iterations=10 ** 9, tolerance=0.0)
        L.lbfgsb200_device_sync()
        t = [time.perf_counter()]
        s = pkg.Solver("rosenbrock", n, p)
        t.append(time.perf_counter())
        s.set_x0(x0)
        t.append(time.perf_counter())
        s.iterate(its)
        t.append(time.perf_counter())
        dev_ms = s.result()["device_ms"]
        s.x(out=out)
        t.append(time.perf_counter())
        s.destroy()
        t.append(time.perf_counter())
        names = ["create", "set_x0", "iterate", "get_x", "destroy"]
        print(json.dumps({"n": n, "buffers": kind, "pass": rep, "iterations": its, "ms": {k: round((t[i + 1] - t[i]) * 1e3, 3) for i, k in enumerate(names)},
                          "iterate_device_ms": round(dev_ms, 3), "total_ms": round((t[-1] - t[0]) * 1e3, 2),
                          "copy_GBps": {"h2d": round(8e-9 * n / max(t[2] - t[1], 1e-9), 1), "d2h": round(8e-9 * n / max(t[4] - t[3], 1e-9), 1)}}), flush=True)
for rep in range(2):
    t0 = time.perf_counter()
    x, info, _ = pkg.solve("rosenbrock", x0g, "wolfe", "par", m=10, max_iterations=its, tolerance=0.0, num_gpus=1)
    dt = time.perf_counter() - t0
    print(json.dumps({"n": n, "buffers": "pageable", "call": "lbfgsb200_solve", "pass": rep, "iterations": int(info["iterations"]),
                      "total_ms": round(dt * 1e3, 2), "device_ms": round(info["device_ms"], 3)}), flush=True)
