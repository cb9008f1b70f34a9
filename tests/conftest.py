"""Shared fixtures.

* ``-m "not gpu"`` : oracle vs the reference's golden vectors, host logic (line-search state
  machines, sharding, x0 generator) and "the C-ABI library loads and exports every symbol" --
  no compute call needs a GPU.
* ``-m gpu``       : the parity tests proper, through the C ABI, against the oracle.

Only tests (and smoke()/bench's CPU-baseline leg) may touch ``oracle/``.
"""
import importlib.util
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_pkg():
    name = "cuda_lbfgs_b200"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(
        name, os.path.join(ROOT, "cuda-lbfgs_b200", "__init__.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def pkg():
    return load_pkg()


@pytest.fixture(scope="session")
def oracle():
    from oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "reference_traces.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def cuda_golden():
    """Outputs of the reference's own CUDA solvers run on a B200 (oracle/make_golden_cuda.py)."""
    with open(os.path.join(ROOT, "tests", "golden", "cuda_reference_traces.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def gpu(pkg):
    """The library on a machine with a device; GPU tests fail (not skip) if the extension is
    missing or no device is visible."""
    L = pkg.lib()
    assert L.lbfgsb200_device_count() >= 1, "gpu-marked test on a machine without a CUDA device"
    return pkg


def unhex(s):
    return float.fromhex(s)


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.maximum(np.abs(b), np.finfo(np.float64).tiny)
    return np.max(np.abs(a - b) / den) if a.size else 0.0


def relvec(a, b):
    """||a-b||_inf / ||b||_inf"""
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
