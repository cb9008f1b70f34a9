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
