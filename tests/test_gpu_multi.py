"""GPU (>= 2 devices): the sharded path -- contiguous shards, one-element halo, packed exchange."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _run(world, *extra):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tests", "multi_gpu_worker.py"), *extra]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
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
    out = _run(min(ndev, 2), *args)
    assert out["ok"], out


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
