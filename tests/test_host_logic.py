"""CPU: host-side logic of the product, no GPU needed.

* the line-search state machines (csrc/ls_logic.h, the code the GPU scalar kernel runs) built for
  the host and compared bit-for-bit with the oracle's restatement of the reference loops;
* the C-ABI library loads and exports every symbol include/lbfgsb200.h declares;
* params defaults, shard arithmetic, the x0 generator; loud failure without a device.
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT

LS = ["backtracking", "interpolation", "wolfe", "backtracking_wolfe"]


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("ls") / "libls_harness.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-std=c++14", "-O2", "-ffp-contract=off", "-fPIC", "-shared",
                           os.path.join(ROOT, "tests", "ls_harness.cpp"), "-o", out])
    L = C.CDLL(out)
    L.harness_ls_poly.restype = C.c_double
    L.harness_ls_poly.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int)]
    for n, k in (("harness_cubic", 6), ("harness_safe_cubic", 6), ("harness_quadratic", 4)):
        getattr(L, n).restype = C.c_double
        getattr(L, n).argtypes = [C.c_double] * k
    return L


def _same(a, b):
    return a == b or (np.isnan(a) and np.isnan(b))


def test_line_search_state_machines_match_oracle(harness, oracle):
    rng = np.random.default_rng(11)
    n_checked = 0
    for trial in range(4000):
        # descent at 0 (c1 < 0), varied curvature so every branch fires: accepts, shrinks,
        # bracket moves, expansions (hi = inf), NaN-producing cubics
        scale = 10.0 ** rng.integers(-3, 4)
        coef = np.array([rng.normal() * scale, -abs(rng.normal()) * scale, rng.normal() * scale * 2,
                         rng.normal() * scale * (trial % 3 == 0), abs(rng.normal()) * scale * (trial % 5 == 0)])
        for kind, ls in enumerate(LS):
            for flavor, fl in enumerate(("seq", "par")):
                want, nf, ng = oracle.ls_poly(ls, fl, coef)
                tr = C.c_int(0)
                got = harness.harness_ls_poly(kind, flavor, coef.ctypes.data_as(C.POINTER(C.c_double)), C.byref(tr))
                assert _same(got, want), (ls, fl, coef, got, want)
                n_checked += 1
    assert n_checked == 4000 * 8


def test_inlined_cuda_line_searches_match_oracle(harness, oracle):
    """FLAVOR_PAR_INLINED (the searches inlined in par/L-BFGS-{Wolfe,Interpolation,Backtracking,Backtracking_Wolfe}.cu):
    step, trial count, success flag and the f(x_host) handed to the next search, decision by decision."""
    harness.harness_ls_poly_inlined.restype = C.c_double
    harness.harness_ls_poly_inlined.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_double, C.c_double,
                                                C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double)]
    rng = np.random.default_rng(23)
    seen_fail = seen_multi = 0
    for trial in range(4000):
        scale = 10.0 ** rng.integers(-3, 4)
        coef = np.array([rng.uniform(-5, 5), -abs(rng.normal()) * scale, rng.normal() * scale,
                         rng.normal() * scale * (trial % 3 == 0), abs(rng.normal()) * scale * (trial % 2 == 0)])
        # the stale f(x_host) is near f(x_k), f(x0) is well above it -- as in a real run
        f_xhost = coef[0] + (rng.normal() * 1e-3 * scale if trial % 4 == 0 else 0.0)
        f_initial = coef[0] + abs(rng.normal()) * 100.0 * scale
        p = coef.ctypes.data_as(C.POINTER(C.c_double))
        for kind, name in enumerate(("backtracking", "interpolation", "wolfe", "backtracking_wolfe")):
            tr, ok, fl = C.c_int(), C.c_int(), C.c_double()
            a = harness.harness_ls_poly_inlined(kind, p, f_xhost, f_initial, C.byref(tr), C.byref(ok), C.byref(fl))
            want, nf, ng, wok, wfl = oracle.ls_poly_inlined(name, coef, f_xhost, f_initial)
            assert _same(a, want), (name, coef, a, want)
            assert ok.value == wok, (name, coef)
            # the oracle counts the extra evaluation at TOL of the bisection search as a trial
            assert tr.value == nf or (name == "backtracking_wolfe" and tr.value + 1 == nf), (name, coef, tr.value, nf)
            assert _same(fl.value, wfl), (name, coef, fl.value, wfl)
            seen_fail += not wok
            seen_multi += nf > 1
    assert seen_fail > 50 and seen_multi > 1000


def test_line_search_trial_counts(harness, oracle):
    # phi(a) = 1 - a + 50 a^2 : steep valley, forces several shrink steps
    coef = np.array([1.0, -1.0, 50.0, 0.0, 0.0])
    p = coef.ctypes.data_as(C.POINTER(C.c_double))
    tr = C.c_int(0)
    a = harness.harness_ls_poly(0, 0, p, C.byref(tr))
    want, nf, ng = oracle.ls_poly("backtracking", "seq", coef)
    assert a == want and nf == 2 * tr.value  # the reference evaluates f(x) AND f(x+ad) per test
    a = harness.harness_ls_poly(2, 1, p, C.byref(tr))
    want, nf, ng = oracle.ls_poly("wolfe", "par", coef)
    assert a == want and nf == tr.value + 1 and ng <= tr.value


def test_interpolation_helpers_match_oracle(harness, oracle):
    rng = np.random.default_rng(5)
    for _ in range(20000):
        a = rng.uniform(-3, 3, 6) * 10.0 ** rng.integers(-2, 3)
        assert _same(harness.harness_cubic(*a), oracle.cubic(*a))
        assert _same(harness.harness_safe_cubic(*a), oracle.safe_cubic(*a))
        assert _same(harness.harness_quadratic(a[0], a[2], a[3], a[4]), oracle.quadratic_interp(a[0], 0.0, a[2], a[3], a[4]))
    # degenerate brackets
    for args in ((0, 0, 1, -1, 1, 1), (1, 0, 2, -1, 3, 1), (0, 1, 1, 0, 1, 0), (0, np.inf, 1, -1, 2, 1)):
        assert _same(harness.harness_safe_cubic(*args), oracle.safe_cubic(*args))
        assert _same(harness.harness_cubic(*args), oracle.cubic(*args))


def test_library_exports_every_declared_symbol(pkg):
    L = pkg.lib()
    header = open(os.path.join(ROOT, "include", "lbfgsb200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = sorted(set(re.findall(r"\b(lbfgsb200_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(L, name), "library does not export " + name
    assert sorted(pkg.EXPORTS) == declared
    assert L.lbfgsb200_version() == 200


def test_library_is_sm100a_native(pkg):
    """The shipped library carries sm_100a SASS for the hot kernels (built by nvcc, not JIT)."""
    pkg.lib()
    r = subprocess.run(["cuobjdump", "-lelf", pkg.LIB_PATH], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in r.stdout
    # Blackwell-native data movement, not a recompiled sm_80 kernel: tensor-map TMA loads (UTMALDG) with mbarrier
    # transaction counting (SYNCS) in pass A / the fused accept kernel, 128-bit global accesses in the streams
    sass = subprocess.run(["cuobjdump", "-sass", pkg.LIB_PATH], capture_output=True, text=True).stdout
    assert sass.count("UTMALDG.2D") >= 10, "tensor-map TMA loads are missing from the SASS"
    assert "SYNCS" in sass and "LDG.E.128" in sass and "STG.E.128" in sass
    for kernel in ("k_accept_gram", "k_combine_trial", "k_gram_tma2d", "k_trial", "k_scalar"):
        assert kernel in sass, kernel


def test_params_defaults_are_the_reference_constants(pkg):
    p = pkg.default_params("seq")
    assert (p.m, p.max_iterations, p.tolerance) == (10, 1000, 1e-5)  # seq/lbfgs.h:22-24
    assert (p.c1, p.c2, p.step0, p.shrink) == (1e-4, 0.9, 1.0, 0.5)   # seq/config.h
    assert (p.backtracking_tol, p.wolfe_min, p.ls_max_trials) == (1e-8, 1e-10, 20)
    assert pkg.default_params("par").c2 == 0.7                        # par/constants.h:6


def test_shard_ranges_cover_and_align(pkg):
    for n in (1, 2, 3, 17, 10000, 10001, 10 ** 8, 2 * 10 ** 9 + 1):
        for P in (1, 2, 3, 4, 8):
            if n < 2 * P and P > 1:
                continue
            pos = 0
            for r in range(P):
                off, ln = pkg.shard_range(n, r, P)
                assert off == pos and ln > 0
                assert off % 2 == 0  # every shard starts on a 16-byte boundary
                pos += ln
            assert pos == n


def test_x0_generator_matches_oracle(pkg, oracle):
    for lo, hi in ((-2, 2), (-1000, 1000), (0.5, 1.5)):
        a = pkg.x0_uniform(1000, lo, hi)
        assert np.array_equal(a, oracle.x0(1000, lo, hi))
        b = pkg.x0_uniform(100, lo, hi, offset=900)
        assert np.array_equal(b, a[900:])  # shards can be generated independently


def test_invalid_arguments_are_rejected(pkg):
    p = pkg.default_params("seq")
    p.line_search = 7  # the reference throws invalid_argument("Unknown line search method")
    with pytest.raises(pkg.LbfgsError, match="Unknown line search method"):
        pkg.Solver("rosenbrock", 100, p)
    p = pkg.default_params("seq", m=0)
    with pytest.raises(pkg.LbfgsError):
        pkg.Solver("rosenbrock", 100, p)


def test_no_cpu_fallback(pkg):
    """Without a CUDA device the solver must fail loudly, never compute on the CPU."""
    if pkg.lib().lbfgsb200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(pkg.LbfgsError, match="no CPU fallback"):
        pkg.solve("quadratic", np.zeros(16))


def test_header_is_plain_c(tmp_path):
    """include/lbfgsb200.h must be consumable by a C compiler (the drop-in boundary is a C ABI)."""
    src = tmp_path / "t.c"
    src.write_text('#include "%s"\nint main(void) { lbfgsb200_params_t p; (void)p; return LBFGSB200_VERSION == 100 ? 0 : 1; }\n'
                   % os.path.join(ROOT, "include", "lbfgsb200.h"))
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.check_call([cc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", str(src)])


def test_compat_shim_host_behaviour(pkg, tmp_path):
    """include/lbfgsb200_compat.hpp without a GPU: objective recognition and the reference's exceptions
    (seq/lbfgs.cpp:69 for an unknown method; a foreign objective is refused instead of run on the CPU)."""
    pkg.lib()
    libdir = os.path.dirname(pkg.LIB_PATH)
    exe = str(tmp_path / "compat_host_check")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    odir = os.path.join(ROOT, "oracle")
    subprocess.check_call(["make", "-s", "-C", odir, "oracle"])
    subprocess.check_call([cxx, "-std=c++14", "-O1", "-Wall", os.path.join(ROOT, "tests", "compat_host_check.cpp"),
                           "-L" + libdir, "-llbfgsb200", "-L" + odir, "-llbfgs_oracle", "-Wl,-rpath," + libdir,
                           "-Wl,-rpath," + odir, "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.strip().splitlines()
    b, e = [i for i, l in enumerate(lines) if l.startswith("cuda_log_begin")][0], [i for i, l in enumerate(lines) if l.startswith("cuda_log_end")][0]
    cuda_log, lines = lines[b + 1:e], lines[:b] + lines[e + 1:]
    got = dict(line.split(": ", 1) for line in lines)
    assert got["quadratic"] == "0" and got["rosenbrock"] == "1" and got["tridiag6"] == "2"
    # the reference's tridiagonal generator assert()s on its dimension (seq/benchmark.cpp:18): probes have that size
    assert "ABORT" not in r.stdout
    assert got["tridiag10000"] == "2" and got["rosenbrock4097"] == "1" and got["quadratic1"] == "0"
    assert got["bound_objective"].startswith("solved 64") or "no usable CUDA device" in got["bound_objective"] \
        or "no CPU fallback" in got["bound_objective"], got["bound_objective"]
    assert got["named_objective"] == "invalid_argument Unknown line search method: newton"
    # LBFGS_CUDA's progress output: same lines as the reference's CUDA solver printed on a B200 for this case
    # (tests/golden/cuda_reference_traces.json), numbers to the 6 significant digits of operator<<
    import json
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "cuda_reference_traces.json")))["traces"]["wolfe_rosen_1e4"]
    want = want["stdout_of_longest_run"].strip().splitlines()
    assert len(cuda_log) == len(want) == 1 + 3 * 20, (len(cuda_log), len(want))
    for g, w in zip(cuda_log, want):
        gl, gv = g.rsplit(" ", 1) if " " in g else (g, None)
        wl, wv = w.rsplit(" ", 1) if " " in w else (w, None)
        assert gl == wl, (g, w)
        if wv is not None:
            assert abs(float(gv) - float(wv)) <= 2e-5 * abs(float(wv)), (g, w)
    assert got["quartic"] == "-1" and got["mismatched_gradient"] == "-1"
    assert got["unknown_method"] == "invalid_argument Unknown line search method: newton"
    assert got["foreign_objective"].startswith("invalid_argument lbfgsb200: the objective is not one of the built-in")
