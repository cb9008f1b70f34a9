#!/usr/bin/env python
"""BASELINE config 5: history-size sweep m = 3..50 at n = 1e8 sharded over the GPUs of one box,
explicit two-loop vs compact form, iterations/s and fraction of the measured HBM peak per m.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29555 benchmarks/msweep_multi.py [--size N] [--hists 3,5,10,20,30,50]

One JSON line per (m, direction) on rank 0.  Timing: CUDA events on each rank's solver stream,
max over ranks, 20 steady-state iterations after m+2 fill iterations.
"""
import argparse
import importlib.util
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_pkg():
    name = "cuda_lbfgs_b200"
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "cuda-lbfgs_b200", "__init__.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=100_000_000)
    ap.add_argument("--hists", default="3,5,10,20,30,50")
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    pkg = load_pkg()
    torch.cuda.set_device(local)
    pkg._check(pkg.lib().lbfgsb200_set_device(local), "set_device")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ids = [pkg.Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    comm = pkg.Comm(ids[0], rank, world)
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    off, ln = pkg.shard_range(a.size, rank, world)
    x0 = pkg.PinnedArray(ln)
    pkg.x0_uniform(ln, -2.0, 2.0, offset=off, out=x0.array)
    for m in [int(v) for v in a.hists.split(",")]:
        for direction in ("two_loop", "compact"):
            p = pkg.default_params("par", line_search="wolfe", m=m, max_iterations=10 ** 9, tolerance=0.0, direction=direction,
                                   use_graph=1)
            s = pkg.Solver("rosenbrock", a.size, p, comm=comm, trace_rows=m + 2 + a.steps + 2)
            s.set_x0(x0.array)
            s.iterate(m + 2)
            dist.barrier()
            s.iterate(a.steps)
            r, tr = s.result(), s.trace()
            s.destroy()
            t = torch.tensor([r["device_ms"]], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            if rank == 0:
                gbs = r["bytes_moved"] / (ms * 1e-3) / 1e9
                print(json.dumps({"config": "5: history sweep, n=%d on %d GPUs, Wolfe" % (a.size, world), "m": m, "direction": direction,
                                  "iterations_per_s": a.steps / (ms * 1e-3), "ms_per_iteration": ms / a.steps,
                                  "GBps_per_gpu": gbs, "frac_of_measured_peak": gbs / peak, "frac_of_8TBps": gbs / 8000.0,
                                  "trials_per_iteration": float(np.mean(tr[-a.steps:, 4])), "f": r["f"]}), flush=True)
    comm.destroy()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
