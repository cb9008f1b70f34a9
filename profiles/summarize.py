#!/usr/bin/env python
"""Turn the raw ncu outputs brought back in gpurun_out/ into the committed summaries:

  profiles/launches_<tag>.csv      (copied)  per-launch device times, cold-cache & serialised
  profiles/launch_shares_<tag>.txt            kernel shares of a steady-state window
  profiles/ncu_kernels_<tag>.txt              key metrics of the --set full captures
  profiles/ncu_traffic.json                   dram bytes per launch per kernel (read by bench.py)

    python profiles/summarize.py <tag> <launches.csv> [<report.ncu-rep>]
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_barrier",
        "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_mio_throttle"]


def short(name):
    return name.split("(")[0].replace("void ", "").replace("lb::", "")


def launches(tag, path):
    shutil.copy(path, os.path.join(HERE, "launches_%s.csv" % tag))
    lines = [l for l in open(path) if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        k = short(row["Kernel Name"])
        tot[k] += v
        cnt[k] += 1
    T = sum(tot.values())
    with open(os.path.join(HERE, "launch_shares_%s.txt" % tag), "w") as fh:
        fh.write("# ncu --metrics gpu__time_duration.sum --clock-control none ; window of %d launches, %.1f us total\n" % (sum(cnt.values()), T))
        fh.write("# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes\n")
        for k in sorted(tot, key=lambda k: -tot[k]):
            fh.write("%-44s launches=%4d total_us=%10.1f avg_us=%9.1f share=%5.1f%%\n" % (k, cnt[k], tot[k], tot[k] / cnt[k], 100 * tot[k] / T))
    print(open(os.path.join(HERE, "launch_shares_%s.txt" % tag)).read())


def full(tag, rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    traffic = {}
    tp = os.path.join(HERE, "ncu_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp))
    agg = collections.defaultdict(list)
    with open(os.path.join(HERE, "ncu_kernels_%s.txt" % tag), "w") as fh:
        fh.write("# ncu --set full --clock-control none --import-source on ; one block per captured launch\n")
        for r in rows[2:]:
            name = short(r[hdr.index("Kernel Name")])
            fh.write("== %s\n" % name)
            for k in KEYS:
                if k in hdr:
                    fh.write("   %-72s %s %s\n" % (k, r[hdr.index(k)], units[hdr.index(k)]))
            def val(key):
                v = float(r[hdr.index(key)].replace(",", ""))
                u = units[hdr.index(key)]
                return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
            agg[name.split("<")[0]].append(val("dram__bytes_read.sum") + val("dram__bytes_write.sum"))
    for k, v in agg.items():
        traffic[k] = {"dram_bytes_per_launch": sum(v) / len(v), "launches_captured": len(v), "source": "profiles/ncu_kernels_%s.txt" % tag}
    json.dump(traffic, open(tp, "w"), indent=1, sort_keys=True)
    print(open(os.path.join(HERE, "ncu_kernels_%s.txt" % tag)).read()[:6000])


if __name__ == "__main__":
    tag = sys.argv[1]
    launches(tag, sys.argv[2])
    if len(sys.argv) > 3:
        full(tag, sys.argv[3])
