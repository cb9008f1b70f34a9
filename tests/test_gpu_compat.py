"""GPU: the reference's own sequential driver (sequential-implementation/main.cpp + benchmark.cpp,
compiled UNMODIFIED by oracle/Makefile) running on the drop-in through the C++ shim
include/lbfgsb200_compat.hpp.  As shipped it minimises the separable quadratic, dim 10 000,
x0 ~ U(-1000,1000), "backtracking", m=10, tol 1e-8: the reference converges at k=2 with
f = 3.44e-23."""
import os
import re
import subprocess

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_unmodified_reference_main_runs_on_the_drop_in(gpu):
    exe = os.path.join(ROOT, "oracle", "_ref", "seq_main_on_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/seq_main_on_b200 not built (needs /root/reference at build time)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Converged!" in r.stdout and "Function: Quadratic Function" in r.stdout
    m = re.search(r"Optimum value: ([-+0-9.eE]+)", r.stdout)
    assert m and abs(float(m.group(1))) < 1e-15, r.stdout


def _cuda_log(text):
    """(alphas, norm_g per iteration, converged-at or None, final "Optimum value") of a CUDA-tree program's stdout."""
    alphas, gnorms, conv, final = [], [], None, None
    for line in text.splitlines():
        if line.startswith("alpha: "):
            alphas.append(float(line.split()[1]))
        elif line.startswith("Iteration ") and "norm_g" in line:
            gnorms.append(float(line.rsplit("=", 1)[1]))
        elif line.startswith("Convergence achieved at iteration"):
            conv = int(line.split()[-1])
        elif line.startswith("Optimum value: "):
            final = float(line.split()[-1])
    return alphas, gnorms, conv, final


@pytest.mark.parametrize("variant", ["host", "wolfe"])
def test_cuda_tree_mains_run_on_the_drop_in(gpu, variant):
    """The main() of parallel-implementation/L-BFGS.cu (n = 5, host Wolfe search) and of L-BFGS-Wolfe.cu (n = 50 000,
    inlined Wolfe, the file BASELINE config 2 is quoted on), cut out of their files at build time and linked against the
    shim (oracle/Makefile, target cudamain), next to the reference program itself (the same main() inside
    oracle/_ref/libref_cuda_<variant>.so, run on this GPU): same output structure, the first iterations print the same
    alpha and |g| to the 6 digits of operator<<, both converge to the mains' tolerance 1e-1."""
    from oracle import CudaRef
    exe = os.path.join(ROOT, "oracle", "_ref", "%s_main_on_b200" % variant)
    if not os.path.exists(exe) or not CudaRef.available(variant):
        pytest.skip("oracle/_ref/%s_main_on_b200 not built (needs /root/reference at build time)" % variant)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    want = CudaRef(variant).own_main()
    ga, gg, gconv, gfinal = _cuda_log(r.stdout)
    wa, wg, wconv, wfinal = _cuda_log(want)
    assert r.stdout.startswith("First x: ") and "Found solution: " in r.stdout
    # the x0 the program printed is the reference's
    assert r.stdout.split("Starting")[0] == want.split("Starting")[0]
    assert gconv is not None and wconv is not None, (gconv, wconv)
    # par/L-BFGS.cu hands the gradient of x0 to every host line search (:199, :293), which the
    # drop-in deliberately does not reproduce: its program agrees with the drop-in only while that has no effect
    head = min(4 if variant == "host" else 15, len(wa), len(ga))
    assert head >= 3
    for k in range(head):
        assert abs(ga[k] - wa[k]) <= 2e-5 * abs(wa[k]), (k, ga[k], wa[k])
        assert abs(gg[k] - wg[k]) <= 2e-5 * abs(wg[k]), (k, gg[k], wg[k])
    assert gg[-1] <= 1e-1 and wg[-1] <= 1e-1
    if variant == "host":  # n = 5: both reach the same minimiser (f* = 0 at x = 1)
        assert gfinal < 1e-2 and wfinal < 1e-2
    else:  # long-horizon trajectories are chaotic at rounding level: same order of magnitude of work
        assert 0.5 <= (gconv + 1) / (wconv + 1) <= 2.0, (gconv, wconv)
    print("%s main(): converged at iteration %d (reference program: %d), final f %.6g (reference %.6g)" %
          (variant, gconv, wconv, gfinal, wfinal))


def test_unmodified_reference_main_on_two_gpus(gpu):
    """LBFGSB200_NUM_GPUS=2: the same unmodified binary, the library shards the solve over two GPUs of this process."""
    exe = os.path.join(ROOT, "oracle", "_ref", "seq_main_on_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/seq_main_on_b200 not built (needs /root/reference at build time)")
    if gpu.lib().lbfgsb200_device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300, env=dict(os.environ, LBFGSB200_NUM_GPUS="2"))
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Converged!" in r.stdout
    m = re.search(r"Optimum value: ([-+0-9.eE]+)", r.stdout)
    assert m and abs(float(m.group(1))) < 1e-15, r.stdout


@pytest.fixture(scope="module")
def verbose_exe(tmp_path_factory, gpu):
    import subprocess as sp
    out = str(tmp_path_factory.mktemp("vm") / "verbose_main")
    libdir = os.path.dirname(gpu.LIB_PATH)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    sp.check_call([cxx, "-std=c++14", "-O1", os.path.join(ROOT, "tests", "verbose_main.cpp"), "-L" + libdir, "-llbfgsb200",
                   "-Wl,-rpath," + libdir, "-o", out])
    return out


def _numbers(line):
    return [float(v) for v in re.findall(r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?", line)]


@pytest.mark.parametrize("case,golden_name", [("rosen5", "verbose_rosen5_backtracking.txt"),
                                              ("tridiag64", "verbose_tridiag64_interpolation.txt")])
def test_verbose_output_matches_the_reference_stdout(verbose_exe, case, golden_name):
    """verbose=true through the shim prints what the reference prints (seq/lbfgs.cpp:76-78, :82, :201):
    same number of "Iteration k, f = ..., |grad| = ..." lines, same status line, same values to the
    6 significant digits of operator<< (values at the cancellation floor excepted)."""
    want = open(os.path.join(ROOT, "tests", "golden", golden_name)).read().strip().splitlines()
    r = subprocess.run([verbose_exe, case], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    got = r.stdout.strip().splitlines()
    assert len(got) == len(want), (got, want)
    f0 = _numbers(want[0])[1]
    for g, w in zip(got, want):
        if not w.startswith("Iteration"):
            assert g == w  # "Converged!" / "Maximum iterations reached"
            continue
        ng, nw = _numbers(g), _numbers(w)
        assert ng[0] == nw[0]
        for a, b in zip(ng[1:], nw[1:]):
            if abs(b) > 1e-9 * f0:
                assert abs(a - b) <= 2e-5 * abs(b), (g, w)
        if all(abs(b) > 1e-9 * f0 for b in nw[1:]):
            assert g == w, (g, w)  # identical text


def test_shim_error_behaviour_matches_the_reference(verbose_exe):
    r = subprocess.run([verbose_exe, "unknown_method"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 4 and "Unknown line search method: newton" in r.stdout  # seq/lbfgs.cpp:69
    r = subprocess.run([verbose_exe, "unknown_objective"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 3 and "no CPU fallback" in r.stdout


def test_plain_c_host_uses_the_abi(gpu, tmp_path):
    """A C99 program (gcc, no C++ / CUDA on the caller's side) drives the library through include/lbfgsb200.h."""
    exe = str(tmp_path / "c_abi_example")
    libdir = os.path.dirname(gpu.LIB_PATH)
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.check_call([cc, "-std=c99", "-Wall", "-Werror", "-pedantic", os.path.join(ROOT, "tests", "c_abi_example.c"),
                           "-L" + libdir, "-llbfgsb200", "-lm", "-Wl,-rpath," + libdir, "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "status=0 (converged) iterations=2" in r.stdout and "Unknown line search method" in r.stdout
