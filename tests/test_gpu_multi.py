"""GPU (>= 2 devices): the sharded path -- contiguous shards, one-element halo, packed exchange."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _run(world, *extra, env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tests", "multi_gpu_worker.py"), *extra]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, **(env or {})))
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and lines, r.stdout[-2000:] + r.stderr[-4000:]
    return json.loads(lines[-1])


@pytest.mark.parametrize("args", [
    ("--obj", "rosenbrock", "--ls", "wolfe", "--flavor", "par", "--size", "100003"),
    ("--obj", "rosenbrock", "--ls", "backtracking", "--flavor", "seq", "--size", "65536"),
    ("--obj", "tridiag", "--ls", "interpolation", "--flavor", "par", "--size", "40001", "--iters", "9"),
    ("--obj", "quadratic", "--ls", "wolfe", "--flavor", "par", "--size", "5000", "--iters", "3"),
])
def test_sharded_equals_single_gpu(gpu, args):
    ndev = gpu.lib().lbfgsb200_device_count()
    if ndev < 2:
        pytest.skip("needs >= 2 GPUs")
    out = _run(min(ndev, 2), *args, "--dir", "two_loop")
    assert out["ok"], out


@pytest.mark.parametrize("graph", [0, 1])
@pytest.mark.parametrize("args", [
    ("--obj", "rosenbrock", "--ls", "wolfe", "--flavor", "par", "--size", "100003"),
    ("--obj", "rosenbrock", "--ls", "interpolation", "--flavor", "par", "--size", "65536", "--hist", "5"),
    ("--obj", "tridiag", "--ls", "backtracking", "--flavor", "seq", "--size", "40001", "--iters", "9"),
    ("--obj", "quadratic", "--ls", "wolfe", "--flavor", "par", "--size", "5000", "--iters", "3"),
])
def test_sharded_fused_compact_flow_equals_single_gpu(gpu, args, graph):
    """The default flow (fused accept + pass A, fused direction + first trial): t + 1 exchanges per iteration through the
    one-way NVLink mailboxes, the neighbours' boundary d formed locally from per-slot boundary values."""
    ndev = gpu.lib().lbfgsb200_device_count()
    if ndev < 2:
        pytest.skip("needs >= 2 GPUs")
    out = _run(min(ndev, 2), *args, "--dir", "compact", "--graph", str(graph))
    assert out["ok"], out


@pytest.mark.parametrize("graph", [0, 1])
@pytest.mark.parametrize("args", [
    ("--ls", "wolfe", "--flavor", "par", "--size", "100003", "--dir", "compact"),
    ("--ls", "interpolation", "--flavor", "par", "--size", "20001", "--dir", "two_loop"),
    ("--ls", "backtracking", "--flavor", "seq", "--size", "4097", "--dir", "compact", "--hist", "5"),
])
def test_user_objective_on_a_sharded_solver(gpu, tmp_path, args, graph):
    """lbfgsb200_create_callback_sharded: a user-written chained Rosenbrock (tests/custom_objective.cu) evaluates its
    shard with the neighbours' boundary values from lbfgsb200_device_halo(); host-stepped and recorded into the graph."""
    ndev = gpu.lib().lbfgsb200_device_count()
    if ndev < 2:
        pytest.skip("needs >= 2 GPUs")
    out_lib = str(tmp_path / "libcustom_objective.so")
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-fmad=false", "-shared",
                           "-Xcompiler", "-fPIC", os.path.join(ROOT, "tests", "custom_objective.cu"), "-o", out_lib])
    out = _run(min(ndev, 2), "--obj", "callback", "--userlib", out_lib, *args, "--graph", str(graph))
    assert out["ok"], out


@pytest.mark.parametrize("args", [
    ("--obj", "rosenbrock", "--ls", "wolfe", "--flavor", "par", "--size", "100003", "--dir", "compact"),
    ("--obj", "rosenbrock", "--ls", "wolfe", "--flavor", "par", "--size", "65536", "--dir", "two_loop"),
    ("--obj", "tridiag", "--ls", "interpolation", "--flavor", "par", "--size", "40001", "--iters", "9", "--dir", "compact"),
])
def test_nccl_exchange_path_equals_single_gpu(gpu, args):
    """LBFGSB200_P2P=0: the packed exchange goes through ncclAllGather instead of the mailboxes (what runs where peer
    access is unavailable); host-stepped loop."""
    ndev = gpu.lib().lbfgsb200_device_count()
    if ndev < 2:
        pytest.skip("needs >= 2 GPUs")
    out = _run(min(ndev, 2), *args, env={"LBFGSB200_P2P": "0"})
    assert out["ok"], out


@pytest.mark.parametrize("objective,ls,flavor,n,K", [("rosenbrock", "wolfe", "par", 1000003, 20), ("tridiag", "interpolation", "par", 250001, 9),
                                                     ("rosenbrock", "backtracking", "seq", 65536, 20)])
def test_one_process_drives_several_gpus_behind_solve(gpu, objective, ls, flavor, n, K):
    """lbfgsb200_solve with params.num_gpus = P: one host call, the library shards over P devices itself (one worker
    thread per GPU, in-process peer-access mailboxes, no NCCL).  Same iterates as the single-GPU call."""
    ndev = gpu.lib().lbfgsb200_device_count()
    if ndev < 2:
        pytest.skip("needs >= 2 GPUs")
    import numpy as np
    x0 = gpu.x0_uniform(n, -2, 2)
    xs, infos, trs = gpu.solve(objective, x0, ls, flavor, trace_rows=K, max_iterations=K, num_gpus=1)
    for P in sorted({2, min(ndev, 8)}):
        x, info, tr = gpu.solve(objective, x0, ls, flavor, trace_rows=K, max_iterations=K, num_gpus=P)
        assert info["status"] == infos["status"] and info["iterations"] == infos["iterations"], (P, info, infos)
        assert np.array_equal(tr[:, 4], trs[:, 4]) and np.array_equal(tr[:, 5], trs[:, 5]), P
        err = float(np.max(np.abs(x - xs)) / np.max(np.abs(xs)))
        assert err <= 1e-10 and abs(info["f"] - infos["f"]) <= 1e-10 * abs(infos["f"]), (P, err)


@pytest.mark.parametrize("args", [
    ("--obj", "rosenbrock", "--ls", "wolfe", "--flavor", "par", "--size", "1000003", "--dir", "compact", "--graph", "1"),
    ("--obj", "rosenbrock", "--ls", "interpolation", "--flavor", "par", "--size", "400009", "--hist", "20", "--dir", "compact"),
    ("--obj", "tridiag", "--ls", "wolfe", "--flavor", "par", "--size", "250001", "--iters", "9"),
])
def test_sharded_equals_single_gpu_on_every_device_of_the_node(gpu, args):
    """The same comparison with one rank per visible device (4 or 8): the mailbox exchange with more than one
    peer, the 3(2m+1)-row Gram exchange, odd shard sizes."""
    ndev = gpu.lib().lbfgsb200_device_count()
    if ndev < 4:
        pytest.skip("needs >= 4 GPUs")
    out = _run(min(ndev, 8), *args)
    assert out["ok"], out
