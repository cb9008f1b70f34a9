"""Worker for the multi-GPU parity test: one process per GPU (torchrun), NCCL.

Solves the same seeded problem sharded over WORLD_SIZE GPUs and, on rank 0, again on one GPU;
the sharded run must reproduce the single-GPU iterates (halo exchange + packed scalar exchange
are exact data movement; only the summation order of the reductions changes).
Prints one JSON line on rank 0 and exits non-zero on mismatch."""
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
    ap.add_argument("--size", dest="n", type=int, default=100003)
    ap.add_argument("--obj", dest="objective", default="rosenbrock")
    ap.add_argument("--ls", default="wolfe")
    ap.add_argument("--flavor", default="par")
    ap.add_argument("--hist", dest="m", type=int, default=10)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--graph", type=int, default=0)
    ap.add_argument("--dir", dest="direction", default="two_loop")
    ap.add_argument("--userlib", default="", help="--obj callback: tests/custom_objective.cu built as a shared library")
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
    lo, hi = (-2.0, 2.0)
    x0 = pkg.x0_uniform(a.n, lo, hi)
    off, ln = pkg.shard_range(a.n, rank, world)
    assert np.array_equal(pkg.x0_uniform(ln, lo, hi, offset=off), x0[off:off + ln])
    p = pkg.default_params(a.flavor, line_search=a.ls, m=a.m, max_iterations=a.iters, use_graph=a.graph,
                           direction=a.direction)
    if a.objective == "callback":
        # a USER objective on a sharded solver: partial sums of this shard, neighbours' boundary values from the
        # solver's device halo block; compared below with the built-in Rosenbrock on one GPU
        import ctypes as C

        class ShardCtx(C.Structure):
            _fields_ = [("tmp", C.c_void_p), ("halo", C.c_void_p), ("n_global", C.c_size_t)]

        user = C.CDLL(a.userlib)
        tmp = pkg.DeviceBuffer(3 * ln)
        ctx = ShardCtx(tmp.ptr, None, a.n)
        s = pkg.Solver("callback", a.n, p, comm=comm, trace_rows=a.iters,
                       callback=C.cast(user.cb_rosenbrock_sharded, C.c_void_p), user=C.addressof(ctx))
        ctx.halo = s.device_halo()
        a.objective = "rosenbrock"
    else:
        s = pkg.Solver(a.objective, a.n, p, comm=comm, trace_rows=a.iters)
    s.set_x0(np.ascontiguousarray(x0[off:off + ln]))
    s.iterate(a.iters)
    x_local = s.x()
    res, tr = s.result(), s.trace()
    s.destroy()
    if a.userlib and res["graph"] != a.graph:
        print("rank %d: graph mode %d, wanted %d" % (rank, res["graph"], a.graph), file=sys.stderr)
        sys.exit(3)
    parts = [None] * world
    dist.all_gather_object(parts, (off, x_local))
    ok = True
    out = {}
    if rank == 0:
        x = np.empty(a.n)
        for o, xl in parts:
            x[o:o + xl.size] = xl
        xs, info, trs = pkg.solve(a.objective, x0, a.ls, a.flavor, trace_rows=a.iters, m=a.m, max_iterations=a.iters,
                                  direction=a.direction)
        err = float(np.max(np.abs(x - xs)) / np.max(np.abs(xs)))
        ferr = abs(res["f"] - info["f"]) / abs(info["f"])
        same_ctl = bool(np.array_equal(tr[:, 4], trs[:, 4]) and np.array_equal(tr[:, 5], trs[:, 5]))
        ok = err <= 1e-10 and ferr <= 1e-10 and same_ctl and res["iterations"] == info["iterations"]
        out = dict(world=world, n=a.n, iters=int(res["iterations"]), max_rel_dx=err, rel_df=ferr, same_trials_history=same_ctl,
                   f=res["f"], f_single=info["f"], ok=ok)
        print(json.dumps(out))
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    comm.destroy()
    dist.destroy_process_group()
    return 0 if int(flag.item()) == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
