"""GPU: the vector_utils / kernel layer through the C ABI, against the oracle.

Element-wise outputs (axpy, scal, gradients, x_new, s, y) must be BIT-IDENTICAL to the oracle:
the kernels are compiled without FMA contraction and keep the reference's operation order.
Reductions differ from the reference's naive left-to-right sums only by summation order; the
tolerance is 1e-13 relative (the tree sum is the more accurate of the two) and every reduction
must be bitwise reproducible run to run (fixed order, no atomics)."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SIZES = [1, 2, 3, 255, 256, 257, 2047, 2048, 2049, 4096, 10000, 65537, 1000003]


def _dev(pkg, a):
    return pkg.DeviceBuffer(a.size, a)


def test_dot_and_norm(gpu, oracle):
    rng = np.random.default_rng(1)
    for n in SIZES:
        a, b = rng.standard_normal(n), rng.standard_normal(n)
        da, db, out = _dev(gpu, a), _dev(gpu, b), gpu.DeviceBuffer(2)
        assert gpu.lib().lbfgsb200_dot(da.ptr, db.ptr, n, out.ptr, None) == 0
        got = out.download(1)[0]
        exact = math.fsum(a * b)
        scale = math.fsum(np.abs(a * b))
        assert abs(got - exact) <= 4e-16 * scale + 1e-300, (n, got, exact)
        assert abs(got - oracle.dot(a, b)) <= 1e-13 * scale
        assert gpu.lib().lbfgsb200_dot(da.ptr, db.ptr, n, out.ptr, None) == 0
        assert out.download(1)[0] == got  # deterministic
        assert gpu.lib().lbfgsb200_nrm2(da.ptr, n, out.ptr, None) == 0
        assert abs(out.download(1)[0] - oracle.norm(a)) <= 1e-13 * oracle.norm(a)


def test_dot_of_ones_is_exact(gpu):
    for n in (1, 12345, 1 << 20, 3 * (1 << 20) + 1):
        a = np.ones(n)
        da, out = _dev(gpu, a), gpu.DeviceBuffer(1)
        gpu.lib().lbfgsb200_dot(da.ptr, da.ptr, n, out.ptr, None)
        assert out.download(1)[0] == float(n)


def test_axpy_scal_bitwise(gpu):
    rng = np.random.default_rng(2)
    for n in SIZES:
        x, y = rng.standard_normal(n), rng.standard_normal(n)
        alpha = np.array([-0.7310585786300049])
        dx, dy, da, dout = _dev(gpu, x), _dev(gpu, y), _dev(gpu, alpha), gpu.DeviceBuffer(n)
        assert gpu.lib().lbfgsb200_axpy(da.ptr, dx.ptr, dy.ptr, n, None) == 0
        assert np.array_equal(dy.download(), y + alpha[0] * x)  # mul, then add: two roundings
        assert gpu.lib().lbfgsb200_scal(da.ptr, dx.ptr, dout.ptr, n, None) == 0
        assert np.array_equal(dout.download(), alpha[0] * x)


@pytest.mark.parametrize("objective", ["quadratic", "rosenbrock", "tridiag"])
def test_eval_trial_matches_oracle(gpu, oracle, objective):
    rng = np.random.default_rng(3)
    for n in SIZES:
        x, d = rng.uniform(-2, 2, n), rng.standard_normal(n)
        for alpha in (0.0, 1.0, 0.37):
            xt = x + alpha * d
            dx, dd, da = _dev(gpu, x), _dev(gpu, d), _dev(gpu, np.array([alpha]))
            g_out, out3 = gpu.DeviceBuffer(n), gpu.DeviceBuffer(3)
            rc = gpu.lib().lbfgsb200_eval_trial(gpu.OBJ[objective], dx.ptr, dd.ptr, da.ptr, n, g_out.ptr, out3.ptr, None)
            assert rc == 0
            g_ref = oracle.grad(objective, xt)
            assert np.array_equal(g_out.download(), g_ref), (objective, n, alpha)  # bit-identical gradient
            f, gd, gg = out3.download()
            f_ref = oracle.f(objective, xt)
            assert abs(f - f_ref) <= 1e-13 * abs(f_ref) + 1e-300, (objective, n, f, f_ref)
            gd_ref, gd_scale = oracle.dot(g_ref, d), math.fsum(np.abs(g_ref * d))
            assert abs(gd - gd_ref) <= 1e-13 * gd_scale + 1e-300
            assert abs(gg - oracle.dot(g_ref, g_ref)) <= 1e-13 * gg + 1e-300
            # without the gradient store (the solver's mode) the sums are the same bits
            rc = gpu.lib().lbfgsb200_eval_trial(gpu.OBJ[objective], dx.ptr, dd.ptr, da.ptr, n, None, out3.ptr, None)
            assert rc == 0 and np.array_equal(out3.download(), [f, gd, gg])


@pytest.mark.parametrize("objective", ["quadratic", "rosenbrock", "tridiag"])
def test_accept_matches_oracle(gpu, oracle, objective):
    rng = np.random.default_rng(4)
    for n in SIZES:
        x, d = rng.uniform(-2, 2, n), rng.standard_normal(n)
        g_old = oracle.grad(objective, x)
        alpha = 0.25
        dx, dd, dg, da = _dev(gpu, x), _dev(gpu, d), _dev(gpu, g_old), _dev(gpu, np.array([alpha]))
        ds, dy, out5 = gpu.DeviceBuffer(n), gpu.DeviceBuffer(n), gpu.DeviceBuffer(5)
        rc = gpu.lib().lbfgsb200_accept(gpu.OBJ[objective], dx.ptr, dd.ptr, dg.ptr, da.ptr, n, ds.ptr, dy.ptr, out5.ptr, None)
        assert rc == 0
        x_new = x + alpha * d
        g_new = oracle.grad(objective, x_new)
        s, y = x_new - x, g_new - g_old
        assert np.array_equal(dx.download(), x_new)
        assert np.array_equal(dg.download(), g_new)
        assert np.array_equal(ds.download(), s) and np.array_equal(dy.download(), y)
        f, gg, sy, yy, sg = out5.download()
        for got, a, b in ((gg, g_new, g_new), (sy, s, y), (yy, y, y), (sg, s, g_new)):
            assert abs(got - oracle.dot(a, b)) <= 1e-13 * math.fsum(np.abs(a * b)) + 1e-300
        f_ref = oracle.f(objective, x_new)
        assert abs(f - f_ref) <= 1e-13 * abs(f_ref) + 1e-300


def test_two_loop_matches_oracle(gpu, oracle):
    rng = np.random.default_rng(5)
    for n in (1, 2, 257, 4096, 10001, 300007):
        for h in (0, 1, 2, 5, 10, 20):
            stride = (n + 31) // 32 * 32
            S = np.zeros((max(h, 1), stride))
            Y = np.zeros((max(h, 1), stride))
            S[:h, :n] = rng.standard_normal((h, n))
            Y[:h, :n] = S[:h, :n] * rng.uniform(0.5, 2.0, (1, n)) + 0.05 * rng.standard_normal((h, n))
            g = rng.standard_normal(n)
            dS, dY, dg = _dev(gpu, S), _dev(gpu, Y), _dev(gpu, g)
            dd, out2 = gpu.DeviceBuffer(n), gpu.DeviceBuffer(2)
            rc = gpu.lib().lbfgsb200_two_loop(dg.ptr, dS.ptr, dY.ptr, h, n, stride, dd.ptr, out2.ptr, None)
            assert rc == 0, gpu.lib().lbfgsb200_last_error()
            d_ref, fell = oracle.two_loop(g, S[:h, :n].copy(), Y[:h, :n].copy()) if h else (-g, False)
            d = dd.download()
            gd, fb = out2.download()
            if h == 0:
                assert np.array_equal(d, -g) and fb == 1.0
                continue
            assert bool(fb) == fell
            err = np.max(np.abs(d - d_ref)) / np.max(np.abs(d_ref))
            assert err <= 1e-12, (n, h, err)
            gd_ref = oracle.dot(g, d_ref)
            assert abs(gd - gd_ref) <= 1e-11 * abs(gd_ref)
            # deterministic
            rc = gpu.lib().lbfgsb200_two_loop(dg.ptr, dS.ptr, dY.ptr, h, n, stride, dd.ptr, out2.ptr, None)
            assert rc == 0 and np.array_equal(dd.download(), d)


def test_two_loop_falls_back_like_the_reference(gpu, oracle):
    """Non-positive gamma / non-finite rho => d = -g (seq/lbfgs.cpp:103-108, :119-124)."""
    n, h = 1000, 3
    rng = np.random.default_rng(6)
    stride = 1024
    S, Y = np.zeros((h, stride)), np.zeros((h, stride))
    S[:, :n] = rng.standard_normal((h, n))
    Y[:, :n] = -S[:, :n]  # s.y < 0 for every pair: gamma < 0
    g = rng.standard_normal(n)
    d_ref, fell = oracle.two_loop(g, S[:, :n].copy(), Y[:, :n].copy())
    assert fell
    dS, dY, dg, dd, out2 = _dev(gpu, S), _dev(gpu, Y), _dev(gpu, g), gpu.DeviceBuffer(n), gpu.DeviceBuffer(2)
    assert gpu.lib().lbfgsb200_two_loop(dg.ptr, dS.ptr, dY.ptr, h, n, stride, dd.ptr, out2.ptr, None) == 0
    assert np.array_equal(dd.download(), -g) and out2.download()[1] == 1.0
