"""GPU: user (device-callback) objectives through lbfgsb200_create_callback -- the replacement for
the reference's host std::function f / grad callbacks (seq/lbfgs.h:18-19).

* a user-written Rosenbrock must reproduce the oracle like the built-in objective does;
* the reference's own (unused) dense SPD fixtures, sequential-implementation/matrices.h, are
  known-answer tests: minimise x^T A x + b^T x and compare with `minimum{N}`."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, relvec

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def userlib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("cb") / "libcustom_objective.so")
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-fmad=false", "-shared",
                           "-Xcompiler", "-fPIC", os.path.join(ROOT, "tests", "custom_objective.cu"), "-o", out])
    return C.CDLL(out)


class DenseCtx(C.Structure):
    _fields_ = [("A", C.c_void_p), ("b", C.c_void_p), ("tmp", C.c_void_p)]


def _fn(lib, name):
    return C.cast(getattr(lib, name), C.c_void_p)


@pytest.mark.parametrize("graph", [1, 0])
@pytest.mark.parametrize("direction", ["two_loop", "compact"])
def test_user_rosenbrock_matches_oracle(gpu, oracle, userlib, direction, graph):
    """graph = 1 (the default): the callback is recorded into the solver's CUDA graph -- it runs on the host once per
    call site, its kernels are replayed from the device for every evaluation."""
    for n, ls, flavor in ((10000, "wolfe", "par"), (4097, "backtracking", "seq"), (501, "interpolation", "par")):
        x0 = oracle.x0(n, -2, 2)
        K = 20
        xo, io, to = oracle.lbfgs("rosenbrock", x0, ls, flavor, 10, K, 1e-5, trace_rows=K)
        tmp = gpu.DeviceBuffer(3 * n)
        p = gpu.default_params(flavor, line_search=ls, m=10, max_iterations=K, direction=direction, use_graph=graph)
        s = gpu.Solver("callback", n, p, trace_rows=K, callback=_fn(userlib, "cb_rosenbrock"), user=tmp.ptr)
        s.set_x0(x0)
        s.iterate(7)          # resumable: several runs of the same recorded graph
        s.iterate(K + 1)
        x, r, tr = s.x(), s.result(), s.trace()
        s.destroy()
        assert r["graph"] == graph
        assert r["iterations"] == io["iterations"] == K
        assert relvec(x, xo) <= 1e-10, (n, ls, relvec(x, xo))
        assert np.array_equal(tr[:, 4], to[:, 4]) and np.array_equal(tr[:, 5], to[:, 5])
        assert abs(r["f"] - io["f"]) <= 1e-10 * abs(io["f"])


def test_callback_that_cannot_be_captured_falls_back_to_the_stepped_loop(gpu, oracle, userlib):
    n, K = 3001, 15
    x0 = oracle.x0(n, -2, 2)
    xo, io, to = oracle.lbfgs("rosenbrock", x0, "wolfe", "par", 10, K, 1e-5, trace_rows=K)
    tmp = gpu.DeviceBuffer(3 * n)
    p = gpu.default_params("par", line_search="wolfe", m=10, max_iterations=K)  # use_graph = 1 by default
    s = gpu.Solver("callback", n, p, trace_rows=K, callback=_fn(userlib, "cb_rosenbrock_syncing"), user=tmp.ptr)
    s.set_x0(x0)
    s.iterate(K)
    x, r = s.x(), s.result()
    s.destroy()
    assert r["graph"] == 0 and r["iterations"] == K
    assert relvec(x, xo) <= 1e-10
    # and the library is still healthy: a capturable callback right after it runs in graph mode
    s = gpu.Solver("callback", n, p, trace_rows=K, callback=_fn(userlib, "cb_rosenbrock"), user=tmp.ptr)
    s.set_x0(x0)
    s.iterate(K)
    x2, r2 = s.x(), s.result()
    s.destroy()
    assert r2["graph"] == 1 and np.array_equal(x, x2)


def test_dense_spd_known_answers_from_reference_fixtures(gpu, userlib):
    cases = json.load(open(os.path.join(ROOT, "tests", "golden", "dense_spd.json")))["cases"]
    for key, case in cases.items():
        n = int(key)
        A = np.array(case["A"]).reshape(n, n)
        b = np.array(case["b"])
        dA, db, tmp = gpu.DeviceBuffer(n * n, A), gpu.DeviceBuffer(n, b), gpu.DeviceBuffer(3 * n)
        ctx = DenseCtx(dA.ptr, db.ptr, tmp.ptr)
        for ls, flavor in (("wolfe", "par"), ("backtracking", "seq"), ("interpolation", "seq")):
            p = gpu.default_params(flavor, line_search=ls, m=10, max_iterations=500, tolerance=1e-6)
            s = gpu.Solver("callback", n, p, callback=_fn(userlib, "cb_dense"), user=C.addressof(ctx))
            s.set_x0(np.zeros(n))
            s.iterate(501)
            x, r = s.x(), s.result()
            s.destroy()
            # converged, or (as the reference would) stopped by a failed line search at the rounding floor
            assert r["status"] in (0, 2), (n, ls, r)
            # stationarity, and the reference's float-precision `minimum{N}` fixture
            assert np.max(np.abs(2 * A @ x + b)) <= 1e-5, (n, ls, r)
            xmin = np.array(case["minimum"])
            assert np.max(np.abs(x - xmin)) <= 2e-4 * max(1.0, np.max(np.abs(xmin))), (n, ls, np.max(np.abs(x - xmin)))


def test_failing_callback_is_reported(gpu, userlib):
    p = gpu.default_params("seq")
    s = gpu.Solver("callback", 100, p, callback=_fn(userlib, "cb_always_fails"), user=None)
    with pytest.raises(gpu.LbfgsError, match="callback failed"):
        s.set_x0(np.zeros(100))
    s.destroy()
