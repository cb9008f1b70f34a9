"""CPU: the reference arm of bench.py (`--impl reference`) prints the contract's JSON line.

The arm times the UNMODIFIED reference (oracle/_ref, built from /root/reference in the build container) or, where that
library does not exist, the C restatement; here it runs on a reduced sample so the test takes seconds."""
import argparse
import importlib.util
import io
import json
import os
from contextlib import redirect_stdout

from conftest import ROOT


def _bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_reference_arm_prints_the_contract_line(monkeypatch):
    bench = _bench()
    monkeypatch.setattr(bench, "REF_ARM_SAMPLE_N", 20000)
    monkeypatch.delenv("RANK", raising=False)
    buf = io.StringIO()
    with redirect_stdout(buf):
        rc = bench.run_reference_arm(argparse.Namespace(steps=2, warmup=12, gpus=1))
    assert rc == 0
    line = json.loads(buf.getvalue().strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "iterations/s" and line["higher_is_better"] is True
    assert line["metric"] == "L-BFGS iterations/sec (FP64) at n=1e8, m=10" and line["dtype"] == "f64"
    assert line["value"] > 0 and abs(line["ms_per_step"] * line["value"] - 1e3) < 1e-6
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] == 1 and cb["value"] == line["value"] and "n=20000" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["sample_n"] == 20000 and "n=20000" in line["config"]["workload"]


def test_other_ranks_of_the_reference_arm_do_nothing(monkeypatch):
    bench = _bench()
    monkeypatch.setenv("RANK", "1")
    buf = io.StringIO()
    with redirect_stdout(buf):
        assert bench.run_reference_arm(argparse.Namespace(steps=2, warmup=12, gpus=2)) == 0
    assert buf.getvalue() == ""
