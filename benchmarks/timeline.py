#!/usr/bin/env python
"""Device-side timeline of one L-BFGS iteration: where does the time between the vector kernels go?

Every scalar kernel stamps (op, %globaltimer at entry, at exit) when LBFGSB200_TIMELINE is set; the time from
one scalar kernel's exit to the next one's entry is a vector kernel plus two launch gaps, the time inside a
scalar kernel is its own work plus (multi-GPU) the wait for the slowest rank in the packet exchange.

    python benchmarks/timeline.py [--size N] [--hist M] [--graph 0|1] [--iters K]
    python -m torch.distributed.run --nproc-per-node 8 ... benchmarks/timeline.py --gpus 8

Prints one JSON line (rank 0): per op the median scalar-kernel time and the median gap that precedes it.
"""
import argparse
import ctypes as C
import importlib.util
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = {-1: "x0", 0: "init", 1: "iter_begin", 2: "sg", 3: "l1", 4: "l2", 5: "ls_init", 6: "ls_step", 7: "accept",
       8: "compact", 9: "compact_dir", 10: "f_init", 11: "f_accept", 12: "f_dir", 13: "f_fix", 14: "f_begin"}


def load_pkg():
    spec = importlib.util.spec_from_file_location("lbfgsb200_pkg", os.path.join(ROOT, "cuda-lbfgs_b200", "__init__.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["lbfgsb200_pkg"] = mod
    spec.loader.exec_module(mod)
    return mod


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--size", type=int, default=100_000_000)
    ap.add_argument("--hist", type=int, default=10)
    ap.add_argument("--graph", type=int, default=1)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--direction", default="compact")
    ap.add_argument("--sub", type=int, default=0, help="1: also record the sub-marks inside OP_F_ACCEPT (they cost ~1 us each)")
    ap.add_argument("--heat-ms", type=float, default=1500.0,
                    help="iterate this long before the measured window: a GPU that has been idle runs its SMs at ~1 GHz for "
                         "the first tens of milliseconds, which inflates the scalar kernels (not the HBM-bound vector kernels)")
    args = ap.parse_args()
    os.environ["LBFGSB200_TIMELINE"] = str(64 * (args.iters + args.hist + 8))
    if args.sub:
        os.environ["LBFGSB200_TIMELINE_SUB"] = "1"
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    pkg = load_pkg()
    pkg.lib().lbfgsb200_set_device(int(os.environ.get("LOCAL_RANK", "0")))
    comm = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
        uid = [pkg.Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        comm = pkg.Comm(uid[0], rank, world)
    off, ln = pkg.shard_range(args.size, rank, world)
    x0 = pkg.x0_uniform(ln, -2.0, 2.0, seed=42, offset=off)
    p = pkg.default_params("par", line_search="wolfe", m=args.hist, max_iterations=10 ** 9, tolerance=0.0,
                           use_graph=args.graph, direction=args.direction)
    s = pkg.Solver("rosenbrock", args.size, p, comm=comm)
    s.set_x0(x0)
    s.iterate(args.hist + 4)                       # fill the history
    # (an iteration count, not a wall-clock loop: every rank must make the same calls)
    heat_iters = max(50, int(args.heat_ms * 1e-3 / (6e-11 * args.size / world)))
    s.iterate(heat_iters)
    if s.result()["status"] != 3:                  # (LBFGSB200_RUNNING) finished while heating: start over, the clocks are up now
        s.set_x0(x0)
        s.iterate(args.hist + 4)
    L = pkg.lib()
    L.lbfgsb200_debug_timeline.restype = C.c_long
    L.lbfgsb200_debug_timeline.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    L.lbfgsb200_debug_timeline(s.h, None, 0, 1)
    s.iterate(args.iters)
    cap = 64 * args.iters
    rows = np.zeros((cap, 3), dtype=np.uint64)
    nrows = L.lbfgsb200_debug_timeline(s.h, rows.ctypes.data, cap, 0)
    s.destroy()
    rows = rows[:nrows]
    ops = rows[:, 0].astype(np.int64)
    t_in = rows[:, 1].astype(np.int64)
    t_out = rows[:, 2].astype(np.int64)
    inside, before = {}, {}
    prev = None  # last row of a scalar kernel proper (sub-marks carry SM cycles in the third column)
    for i in range(nrows):
        if ops[i] >= 100:
            continue
        if prev is not None:
            name = OPS.get(int(ops[i]), str(int(ops[i])))
            inside.setdefault(name, []).append((t_out[i] - t_in[i]) / 1e3)
            before.setdefault(name, []).append((t_in[i] - t_out[prev]) / 1e3)
        prev = i
    # diagnostic sub-marks (op >= 100) inside OP_F_ACCEPT: ns between consecutive marks of an iteration
    sub = {}
    for i in range(1, nrows):
        if ops[i] >= 100 and ops[i - 1] >= 100:
            sub.setdefault("%d->%d" % (ops[i - 1], ops[i]), []).append(int(t_in[i] - t_in[i - 1]))
    cyc = {}
    for i in range(1, nrows):
        if ops[i] >= 100 and ops[i - 1] >= 100:
            cyc.setdefault("%d->%d" % (ops[i - 1], ops[i]), []).append(int(t_out[i] - t_out[i - 1]))
    if sub and rank == 0:
        print(json.dumps({"ns_between_marks": {k: float(np.median(v)) for k, v in sub.items()},
                          "sm_cycles_between_marks": {k: float(np.median(v)) for k, v in cyc.items()}}), file=sys.stderr)
    main = [i for i in range(nrows) if ops[i] < 100]
    total_us = (t_out[main[-1]] - t_out[main[0]]) / 1e3 if len(main) > 1 else 0.0
    out = {"n_gpus": world, "n": args.size, "m": args.hist, "graph": args.graph, "iters": args.iters,
           "us_per_iteration": total_us / args.iters,
           "scalar_kernel_us_per_iteration": sum(sum(v) for v in inside.values()) / args.iters,
           "ops": {k: {"per_iteration": len(inside[k]) / args.iters, "median_inside_us": float(np.median(inside[k])),
                       "median_gap_before_us": float(np.median(before[k]))} for k in inside}}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        comm.destroy()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
