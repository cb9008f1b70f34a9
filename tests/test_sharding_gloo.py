"""CPU, world_size 2, gloo: the host-side logic of the N>1 path.

What is exercised here (no GPU): the shard arithmetic of the C ABI, per-shard x0 generation,
the out-of-band unique-id exchange the launcher performs, and -- with a numpy model of what the
kernels do -- that the decomposition the GPU path uses (owner-computes f-terms, one-element halo
of x and d refreshed once per outer iteration, per-rank packets summed in rank order) reproduces
the unsharded oracle exactly.  The numpy model lives in this test; the product never runs it."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_pkg

sys.path.insert(0, os.path.join(ROOT, "oracle"))


def _shard_eval(objective, xt_local, left, right, goff, nglob):
    """f-terms owned by this shard + its gradient entries, given the neighbours' boundary values."""
    n = xt_local.size
    ext = np.concatenate([[left if left is not None else 0.0], xt_local, [right if right is not None else 0.0]])
    f = 0.0
    g = np.zeros(n)
    for i in range(n):
        G = goff + i
        l, c, r = ext[i], ext[i + 1], ext[i + 2]
        hl, hr = G > 0, G < nglob - 1
        if objective == "rosenbrock":
            bl, bc = c - l * l, r - c * c
            if hr:
                f += 100.0 * bc * bc + (1 - c) * (1 - c)
            fl = 200.0 * bl if hl else 0.0
            g[i] = fl + (2.0 * (c - 1) - 400.0 * c * bc) if hr else fl
        else:  # tridiag
            f += 1000.0 * c * c + (100.0 * c * r if hr else 0.0)
            gv = 2000.0 * c
            if hl:
                gv = gv + 100.0 * l
            if hr:
                gv = gv + 100.0 * r
            g[i] = gv
    return f, g


def _worker(rank, world, port, n, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = load_pkg()
    from oracle import Oracle
    orc = Oracle()
    # 1. the launcher's unique-id broadcast (bytes from rank 0 reach every rank unchanged)
    ids = [os.urandom(128) if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    assert isinstance(ids[0], bytes) and len(ids[0]) == 128
    # 2. shard arithmetic + per-shard x0 generation
    off, ln = pkg.shard_range(n, rank, world)
    x0 = pkg.x0_uniform(n, -2, 2)
    x_loc = pkg.x0_uniform(ln, -2, 2, offset=off)
    assert np.array_equal(x_loc, x0[off:off + ln])
    rng = np.random.default_rng(123)
    d = rng.standard_normal(n)
    d_loc = d[off:off + ln]
    for objective in ("rosenbrock", "tridiag"):
        for alpha in (0.0, 0.5):
            # 3. packet exchange: [f_partial, x_first, x_last, d_first, d_last] per rank
            pkt = [None] * world
            dist.all_gather_object(pkt, (x_loc[0], x_loc[-1], d_loc[0], d_loc[-1]))
            left = pkt[rank - 1][1] + alpha * pkt[rank - 1][3] if rank > 0 else None
            right = pkt[rank + 1][0] + alpha * pkt[rank + 1][2] if rank < world - 1 else None
            f_loc, g_loc = _shard_eval(objective, x_loc + alpha * d_loc, left, right, off, n)
            parts = [None] * world
            dist.all_gather_object(parts, (f_loc, g_loc))
            f_tot = 0.0
            for fr, _ in parts:  # rank order => identical bits on every rank
                f_tot += fr
            g_all = np.concatenate([gr for _, gr in parts])
            xt = x0 + alpha * d
            assert np.array_equal(g_all, orc.grad(objective, xt)), (objective, alpha)
            assert abs(f_tot - orc.f(objective, xt)) <= 1e-12 * abs(orc.f(objective, xt))
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(tmp, "ok%d" % rank), "w").write("ok")


@pytest.mark.parametrize("n", [2001, 4096])
def test_two_rank_decomposition_matches_oracle(tmp_path, n):
    world = 2
    port = 29500 + (os.getpid() % 400) + (n % 7)
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok%d" % r)) for r in range(world))
