#!/usr/bin/env python
"""bench.py -- L-BFGS iterations/s (FP64) at n=1e8, m=10, Wolfe line search (BASELINE.json config 2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config 2|4|5]

A "step" is one L-BFGS iteration of the Rosenbrock objective (direction from the full history, a Wolfe line search
with fused trial evaluations, the accept/update pass), x0 ~ U(-2,2) drawn like the reference mains (mt19937(42)).
Inputs are synthetic; every vector is 0.8 GB per GPU at N=1, far larger than the 126 MB L2, so nothing survives in
cache between passes.  The solver runs in its default configuration: compact direction in the fused two-kernel flow
(k_accept_gram, k_combine_trial + k_trial for further trials), whole loop as one CUDA graph.

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline     -- the dominant streaming kernel of the step (plus a table of all of them): achieved GB/s = algorithmic
                  bytes per launch / CUDA-event time of every launch in a separate instrumented (host-stepped) run of
                  the same iterations, against MEASURED_PEAKS.json:hbm_gbs.  `traffic` is NOT measured in this run:
                  it is the DRAM-bytes / algorithmic-bytes ratio of the committed ncu capture (profiles/ncu_traffic.json)
                  times the algorithmic bytes, and `traffic_source` says so.
  sustained    -- the same loop for >= 3 s with the clock / power sampler running (power capping shows up here)
  variants     -- the same measurement with the explicit two-loop recursion (host of the reference's algorithm)
  e2e          -- same metric through the host-buffer API from PINNED buffers: create + H2D x0 + iterations + D2H x
  e2e_pageable -- (N=1) the drop-in call itself: lbfgsb200_solve() from pageable numpy buffers
  cpu_baseline -- the UNMODIFIED reference (oracle/_ref: sequential outer loop + the CUDA tree's Wolfe search) on one
                  host core, on a bounded sample
  reference_cuda_on_this_gpu -- the reference's own CUDA solver (parallel-implementation/L-BFGS-Wolfe.cu, unmodified,
                  sm_100) on the same GPU and sample: min and spread of 3 repeats

--config 4: Rosenbrock n=2e9, m=20 sharded over the N GPUs (rank 0 then measures the largest n that fits ONE GPU at
            m=20 as the comparator -- no extrapolation from another size).
--config 5: history sweep m = 3..50 at n=1e8 on the N GPUs, two-loop vs compact, roofline fraction per m.
--impl reference times only the reference's CPU implementation (rank 0; other ranks exit 0).
"""
import argparse
import importlib.util
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))

N_GLOBAL = 100_000_000
M = 10
LINE_SEARCH = "wolfe"
FLAVOR = "par"  # C2 = 0.7 + safeguarded cubic: the CUDA tree's Wolfe search (BASELINE config 2)
OBJECTIVE = "rosenbrock"
CPU_SAMPLE_N = 10_000_000     # cpu_baseline key of the main arm: the size the reference arm uses (its per-element cost still grows
                              # between n = 4e6 and 1e7: 1.15e-7 -> 2.0e-7 s per element and iteration on the B200 host), ~50 s of CPU work
CUDA_REF_SAMPLE_N = 1_000_000  # reference_cuda_on_this_gpu: its iteration is PCIe / host bound, not cache bound
REF_ARM_SAMPLE_N = 10_000_000  # --impl reference: measured directly at n=1e7 with a full history


def load_pkg():
    name = "cuda_lbfgs_b200"
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "cuda-lbfgs_b200", "__init__.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        def num(v):
            try:
                return float(v)
            except ValueError:
                return None
        sm = [num(s[0]) for s in self.samples if s and num(s[0]) is not None]
        mx = [num(s[1]) for s in self.samples if len(s) > 1 and num(s[1]) is not None]
        pw = [num(s[2]) for s in self.samples if len(s) > 2 and num(s[2]) is not None]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples if len(s) >= 7 for i in range(4)
                          if s[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(self.samples)}


def cpu_reference_rate(steps, warmup, n_sample):
    """Steady-state seconds per iteration of the unmodified reference (hybrid: seq outer loop +
    par/line_search.cpp Wolfe), one core, on n_sample elements; scaled linearly to N_GLOBAL."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as om
    kind = "reference"
    if om.Ref.available("par"):
        ref = om.Ref("par")
        x0 = ref.x0(n_sample, -2, 2)

        def run(iters):
            _, info = ref.lbfgs(OBJECTIVE, x0, LINE_SEARCH, M, iters, 1e-5)
            return info["seconds"]
    else:  # oracle/_ref only exists if it was built where /root/reference is mounted
        kind = "port"
        orc = om.Oracle()
        x0 = orc.x0(n_sample, -2, 2)

        def run(iters):
            t = time.perf_counter()
            orc.lbfgs(OBJECTIVE, x0, LINE_SEARCH, FLAVOR, M, iters, 1e-5)
            return time.perf_counter() - t
    warm = max(warmup, 1)
    t_warm = run(warm)
    t_all = run(warm + steps)
    per_iter = max(t_all - t_warm, 1e-9) / steps
    its_sample = 1.0 / per_iter
    return {"value": its_sample * n_sample / N_GLOBAL, "unit": "iterations/s", "cores": 1, "kind": kind,
            "sample": "n=%d (1/%d of the workload), %d steady-state iterations (history full) after %d warm-up "
                      "iterations, measured %.3f s/iteration on one core, scaled linearly in n to n=%d (the reference's "
                      "cost grows faster than linearly in n, so this favours the CPU)" %
                      (n_sample, N_GLOBAL // n_sample, steps, warm, per_iter, N_GLOBAL),
            "host_cores_available": os.cpu_count(), "seconds_per_iteration_at_sample": per_iter,
            "measured_seconds": t_warm + t_all}


def cuda_reference_rate(steps, warmup, n_sample=CUDA_REF_SAMPLE_N, repeats=3):
    """The reference's own CUDA solver for this configuration (parallel-implementation/L-BFGS-Wolfe.cu, unmodified,
    cross-compiled for sm_100: oracle/_ref/libref_cuda_wolfe.so) on the SAME B200: steady-state seconds per
    iteration on n_sample elements, scaled linearly to N_GLOBAL; min and spread over `repeats` measurements.  Its
    iteration ships x and the gradient across PCIe and evaluates f / grad on one host core.
    None when the library was not built (it needs /root/reference at build time)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as om
    if OBJECTIVE != "rosenbrock" or LINE_SEARCH != "wolfe" or not om.CudaRef.available("wolfe"):
        return None
    ref = om.CudaRef("wolfe")
    x0 = om.Oracle().x0(n_sample, -2, 2)

    def run(iters):
        t = time.perf_counter()
        ref.lbfgs(OBJECTIVE, x0, "wolfe", M, iters, 0.0)
        return time.perf_counter() - t
    warm = max(warmup, 1)
    run(1)  # CUDA context / cuBLAS initialisation of that library
    per_iter, discarded = [], []
    for _ in range(repeats):
        t_warm = run(warm)
        t_all = run(warm + steps)
        d = (t_all - t_warm) / steps
        # a repeat whose longer run is not slower than the shorter one by at least a fifth of the shorter run's own
        # per-iteration time did not run the extra iterations (its search broke off) or was disturbed: not a measurement
        (per_iter if d > 0.2 * t_warm / warm else discarded).append(d)
    if not per_iter:
        return {"value": None, "unit": "iterations/s", "kind": "reference (CUDA tree)", "discarded_repeats": discarded,
                "note": "no repeat produced a usable steady-state difference"}
    per_iter.sort()
    med = per_iter[len(per_iter) // 2]
    return {"value": (1.0 / med) * n_sample / N_GLOBAL, "unit": "iterations/s", "kind": "reference (CUDA tree)",
            "source": "parallel-implementation/L-BFGS-Wolfe.cu, unmodified, nvcc defaults, sm_100, cuBLAS; 1 GPU + 1 host core",
            "sample": "n=%d (1/%d of the workload), %d steady-state iterations after %d warm-up, %d repeats, median of the %d "
                      "usable ones scaled linearly in n to n=%d" % (n_sample, N_GLOBAL // n_sample, steps, warm, repeats, len(per_iter), N_GLOBAL),
            "seconds_per_iteration_at_sample": {"median": med, "min": per_iter[0], "max": per_iter[-1], "all": per_iter},
            "discarded_repeats": discarded,
            "value_range": [(1.0 / per_iter[-1]) * n_sample / N_GLOBAL, (1.0 / per_iter[0]) * n_sample / N_GLOBAL]}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # measured directly at n=1e7 (10 % of the workload, every vector far beyond the host caches), history full:
    # 12 warm-up + min(K, 8) timed iterations ~ 1.5 min on one core
    steps = max(1, min(args.steps, 8))
    warm = max(M + 2, min(args.warmup, 12))
    cb = cpu_reference_rate(steps, warm, REF_ARM_SAMPLE_N)
    # BASELINE config 1 (the reference's own CPU case) end to end: Rosenbrock n=1e4, m=10, backtracking,
    # tol 1e-5, unmodified sequential-implementation on one core
    config1 = None
    try:
        import oracle as om
        if om.Ref.available("seq"):
            ref = om.Ref("seq")
            x0 = ref.x0(10000, -2, 2)
            xr, ir = ref.lbfgs("rosenbrock", x0, "backtracking", 10, 20000, 1e-5)
            config1 = {"seconds": ir["seconds"], "iterations": ir["g_evals"] - 1, "status": ir["status"],
                       "iterations_per_s": (ir["g_evals"] - 1) / ir["seconds"], "f": ref.f("rosenbrock", xr)}
    except Exception as e:  # the headline line must still be printed
        config1 = {"error": str(e)}
    line = {"impl": "reference", "metric": "L-BFGS iterations/sec (FP64) at n=1e8, m=10", "value": cb["value"],
            "unit": "iterations/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": 1e3 / cb["value"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "Rosenbrock n=1e8, m=10, Wolfe line search (hybrid CPU reference: sequential-implementation/"
                                   "lbfgs.cpp + parallel-implementation/line_search.cpp, unmodified, 1 core); each step is "
                                   "measured on a bounded sample: n=%d (1/%d of the workload) with a full history, "
                                   "%.2f s per iteration measured, value = that rate x %d / %d" %
                                   (REF_ARM_SAMPLE_N, N_GLOBAL // REF_ARM_SAMPLE_N, cb["seconds_per_iteration_at_sample"],
                                    REF_ARM_SAMPLE_N, N_GLOBAL),
                       "sample_n": REF_ARM_SAMPLE_N, "same_config": False},
            "cpu_baseline": cb, "config1": config1,
            "e2e": {"value": cb["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


class Bench:
    """Shared plumbing of the three configurations: device selection, communicator, barriers."""

    def __init__(self, args):
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.pkg = load_pkg()
        self.L = self.pkg.lib()
        if self.L.lbfgsb200_device_count() < 1:
            raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback")
        self.pkg._check(self.L.lbfgsb200_set_device(self.local_rank), "set_device")
        self.dist = None
        self.comm = None
        if self.world > 1:
            import torch
            import torch.distributed as dist
            torch.cuda.set_device(self.local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            ids = [self.pkg.Comm.unique_id() if self.rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            self.comm = self.pkg.Comm(ids[0], self.rank, self.world)
            self.dist = dist
        self.peak, self.peak_src = measured_peak()
        self.traffic_db = {}
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            try:
                self.traffic_db = json.load(open(tpath))
            except Exception:
                self.traffic_db = {}

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.L.lbfgsb200_device_sync()

    def max_over_ranks(self, v):
        if self.dist is None:
            return v
        import torch
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.comm is not None:
            self.comm.destroy()
        if self.dist is not None:
            self.dist.destroy_process_group()

    def shard_x0(self, n_global):
        off, n_local = self.pkg.shard_range(n_global, self.rank, self.world)
        x0 = self.pkg.PinnedArray(n_local)
        self.pkg.x0_uniform(n_local, -2.0, 2.0, seed=42, offset=off, out=x0.array)
        self.barrier()  # ranks generate shards of different offsets (mt19937 skip-ahead): line up before any exchange
        return x0, n_local

    def kernel_table(self, flow, m, V):
        """algorithmic bytes per launch of each streaming-kernel class, steady state h = m;
        flow = lbfgsb200_result_t.flow of the solver that ran (0 two-loop, 1 compact unfused, 2 compact fused)"""
        h = m
        if flow == 2:
            # k_accept_gram: reads the 2(h-1) kept history rows + x, d, g_old, writes x, g, s, y
            # k_combine_trial: reads the 2h+1 basis vectors + x, writes d (the first line-search trial rides on it)
            return {"gram_rows": ("k_accept_gram", (2 * (h - 1) + 7) * V), "combine": ("k_combine_trial", (2 * h + 3) * V),
                    "trial": ("k_trial", 2 * V)}
        if flow == 1:
            return {"gram_rows": ("k_gram_tma2d", (2 * h + 1) * V), "combine": ("k_combine", (2 * h + 2) * V),
                    "trial": ("k_trial", 2 * V), "accept": ("k_accept", 7 * V)}
        return {"two_loop_pass": ("k_two_loop_pass", (8 * h - 1) * V / (2 * h)), "trial": ("k_trial", 2 * V), "accept": ("k_accept", 7 * V)}

    def measure(self, n_global, m, direction, x0, K, W, sample_clocks=False, sustain_s=0.0, graph=1, profile_kernels=True):
        """Device-resident timed region of K iterations after FILL + W warm-up iterations (history full), an optional
        sustained leg, then a separate instrumented run of K more iterations with a CUDA-event pair around every
        streaming kernel."""
        pkg = self.pkg
        FILL = max(0, m + 2 - W)
        V = 8.0 * x0.n
        prm = pkg.default_params(FLAVOR, line_search=LINE_SEARCH, m=m, max_iterations=10 ** 9, tolerance=0.0,
                                 use_graph=graph, direction=direction)
        solver = pkg.Solver(OBJECTIVE, n_global, prm, comm=self.comm, trace_rows=FILL + W + 2 * K + 8)
        solver.set_x0(x0.array)
        if FILL:
            solver.iterate(FILL)  # fill the (s, y) history so every timed step uses all m pairs
        solver.iterate(W)  # warm-up
        launches0 = solver.result()["kernel_launches"]
        sampler = ClockSampler(self.local_rank) if (self.rank == 0 and sample_clocks) else None
        if sampler:
            sampler.start()
        self.barrier()
        t0 = time.perf_counter()
        solver.iterate(K)
        self.barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        if sampler:
            sampler.stop_flag = True
        res = solver.result()
        dev_ms = self.max_over_ranks(res["device_ms"])
        wall_ms = self.max_over_ranks(wall_ms)
        launches = res["kernel_launches"] - launches0
        rows = solver.trace()[FILL + W:FILL + W + K]
        trials = float(np.sum(rows[:, 4])) if len(rows) else 0.0
        bytes_step = res["bytes_moved"] / K  # local shard, algorithmic
        out = {"value": K / (dev_ms / 1e3), "ms_per_step": dev_ms / K, "wall_ms_per_step": wall_ms / K,
               "gpu_launches": int(launches), "trials_per_step": trials / K if K else None, "flow": int(res["flow"]),
               "clocks": sampler.summary() if sampler else None}
        if sustain_s > 0:
            S = max(K, int(math.ceil(sustain_s / (dev_ms / K / 1e3))))
            samp2 = ClockSampler(self.local_rank) if self.rank == 0 else None
            if samp2:
                samp2.start()
            self.barrier()
            solver.iterate(S)
            self.barrier()
            if samp2:
                samp2.stop_flag = True
            r2 = solver.result()
            ms2 = self.max_over_ranks(r2["device_ms"])
            out["sustained"] = {"iterations": S, "seconds": ms2 / 1e3, "value": S / (ms2 / 1e3), "unit": "iterations/s",
                                "ms_per_step": ms2 / S, "vs_short_run": (S / (ms2 / 1e3)) / out["value"],
                                "clocks": samp2.summary() if samp2 else None}
        step_gbs = bytes_step / (dev_ms / K * 1e-3) / 1e9
        out["whole_step"] = {"algorithmic_GB_per_step": bytes_step / 1e9, "achieved_GBps": step_gbs,
                             "frac_of_peak": step_gbs / self.peak, "frac_of_8TBps_nominal": step_gbs / 8000.0}
        if profile_kernels:
            _, classes = solver.iterate_profiled(K)
            res_p = solver.result()
            if self.rank == 0:
                sys.stderr.write("kernel classes (%s): %s\n" % (direction, json.dumps(classes)))
            table = self.kernel_table(res_p["flow"], m, V)
            kernels = {}
            for name, c in classes.items():
                if name in table and c["launches"] > 0 and c["ms"] > 0:
                    kname, nbytes = table[name]
                    ms = c["ms"] / c["launches"]
                    gbs = nbytes / (ms * 1e-3) / 1e9
                    ratio = self.traffic_db.get(kname, {}).get("dram_over_algorithmic")
                    kernels[kname] = {"launches_per_step": c["launches"] / K, "avg_launch_ms": ms,
                                      "algorithmic_bytes_per_launch": nbytes, "achieved_GBps": gbs,
                                      "frac_of_peak": gbs / self.peak, "ms_per_step": c["ms"] / K,
                                      "traffic": ratio * nbytes if ratio else None}
            dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
            out["roofline"] = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_GBps"], "peak": self.peak,
                               "unit": "GB/s", "frac": kernels[dom]["frac_of_peak"], "traffic": kernels[dom]["traffic"],
                               "traffic_source": "committed ncu capture (profiles/ncu_traffic.json: dram bytes / algorithmic "
                                                 "bytes of this kernel) x algorithmic bytes; not measured in this run",
                               "peak_source": self.peak_src,
                               "algorithmic_bytes_per_launch": kernels[dom]["algorithmic_bytes_per_launch"],
                               "avg_launch_ms": kernels[dom]["avg_launch_ms"], "kernels": kernels,
                               "whole_step": out["whole_step"], "instrumented_ms_per_step": res_p["device_ms"] / K}
            out["final"] = {"f": res_p["f"], "gnorm": res_p["gnorm"], "iterations": res_p["iterations"]}
        else:
            out["final"] = {"f": res["f"], "gnorm": res["gnorm"], "iterations": res["iterations"]}
        solver.destroy()
        return out


def run_config2(b, args):
    """The headline: Rosenbrock n (default 1e8), m (default 10), Wolfe, strong scaling over the N GPUs."""
    pkg, rank, world = b.pkg, b.rank, b.world
    n_global, m = args.size, args.hist
    K, W = args.steps, max(args.warmup, 3)
    FILL = max(0, m + 2 - W)
    x0, n_local = b.shard_x0(n_global)
    out_pinned = pkg.PinnedArray(n_local)
    V = 8.0 * n_local
    main_run = b.measure(n_global, m, args.direction, x0, K, W, sample_clocks=True, sustain_s=args.sustain, graph=args.graph)
    other = "two_loop" if args.direction == "compact" else "compact"
    other_run = b.measure(n_global, m, other, x0, K, W, graph=args.graph) if not args.single_variant else None
    params = pkg.default_params(FLAVOR, line_search=LINE_SEARCH, m=m, max_iterations=10 ** 9, tolerance=0.0,
                                use_graph=args.graph, direction=args.direction)

    # ---------------- end to end through the host-buffer API (pinned buffers) ----------------
    its = FILL + W + K
    e2e_runs = []
    for _ in range(2):  # the first call pays one-off costs of the process (staging buffers, function attributes)
        b.barrier()
        t0 = time.perf_counter()
        s2 = pkg.Solver(OBJECTIVE, n_global, params, comm=b.comm, trace_rows=0)
        s2.set_x0(x0.array)             # H2D of x0 from pinned host memory
        s2.iterate(its)
        s2.x(out=out_pinned.array)      # D2H of the result
        b.barrier()
        e2e_runs.append(b.max_over_ranks(time.perf_counter() - t0))
        s2.destroy()
    e2e_s = e2e_runs[-1]
    e2e = {"value": its / e2e_s, "unit": "iterations/s", "h2d_bytes_per_step": 8.0 * n_global / its,
           "d2h_bytes_per_step": 8.0 * n_global / its, "seconds": e2e_s, "first_call_seconds": e2e_runs[0],
           "what": "create + H2D x0 (pinned) + %d iterations from a cold history + D2H x; wall clock, max over ranks; "
                   "byte counts are the one-off 8n-byte copies amortised over the iterations of the call" % its}

    e2e_pageable = None
    cpu_baseline = None
    cuda_reference = None
    if rank == 0 and world == 1:
        # the drop-in call itself: lbfgsb200_solve() from PAGEABLE buffers (what LBFGS() of the shim hands over)
        try:
            xp = np.array(x0.array)  # pageable copy
            pg = []
            for _ in range(2):
                t0 = time.perf_counter()
                _, info, _ = pkg.solve(OBJECTIVE, xp, LINE_SEARCH, FLAVOR, m=m, max_iterations=its, tolerance=0.0,
                                       direction=args.direction, use_graph=args.graph, num_gpus=1)
                pg.append(time.perf_counter() - t0)
            e2e_pageable = {"value": its / pg[-1], "unit": "iterations/s", "seconds": pg[-1], "first_call_seconds": pg[0],
                            "vs_pinned": (its / pg[-1]) / e2e["value"], "iterations": int(info["iterations"]),
                            "what": "lbfgsb200_solve(): create + staged H2D of a pageable x0 + %d iterations + staged D2H "
                                    "into a pageable buffer + destroy; wall clock" % its}
            del xp
        except Exception as e:
            e2e_pageable = {"unavailable": repr(e)}
        if not args.no_cpu_baseline:
            cpu_baseline = cpu_reference_rate(min(K, 3), M + 1, CPU_SAMPLE_N)
            try:
                cuda_reference = cuda_reference_rate(min(K, 20), 12)
            except Exception as e:  # the comparison is informative, never fatal to the bench line
                cuda_reference = {"unavailable": repr(e)}

    if rank == 0:
        metric = "L-BFGS iterations/sec (FP64) at n=1e8, m=10" if (n_global == N_GLOBAL and m == 10) else \
            "L-BFGS iterations/sec (FP64) at n=%.0e, m=%d" % (n_global, m)
        line = {"metric": metric, "value": main_run["value"], "unit": "iterations/s",
                "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": main_run["ms_per_step"], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "Rosenbrock n=%d, m=%d, Wolfe line search (C2=0.7, safeguarded cubic), "
                                       "x0~U(-2,2) mt19937(42), %s direction%s, %s" %
                                       (n_global, m, args.direction,
                                        " (fused accept + pass A, direction + first trial)" if args.direction == "compact" and os.environ.get("LBFGSB200_FUSED", "1") != "0" else "",
                                        "CUDA-graph loop" if args.graph else "host-stepped loop"),
                           "parallelism": "contiguous shards x%d, 1-element halo + packed exchange per sync (NVLink mailboxes)" % world
                                          if world > 1 else "single GPU",
                           "cache": "inputs larger than L2 (each of the 2m+6 vectors is %.2f GB per GPU)" % (V / 1e9),
                           "trials_per_step": main_run["trials_per_step"],
                           "history_fill_iterations_before_warmup": FILL},
                "wall_ms_per_step": main_run["wall_ms_per_step"], "gpu_launches": main_run["gpu_launches"],
                "clocks": main_run["clocks"], "roofline": main_run["roofline"], "sustained": main_run.get("sustained"),
                "e2e": e2e, "e2e_pageable": e2e_pageable, "cpu_baseline": cpu_baseline,
                "reference_cuda_on_this_gpu": cuda_reference, "final": main_run["final"]}
        if other_run is not None:
            line["variants"] = {other: {k: other_run[k] for k in ("value", "ms_per_step", "gpu_launches", "trials_per_step", "final")}}
            line["variants"][other]["roofline"] = {k: other_run["roofline"][k] for k in ("kernel", "achieved", "frac", "whole_step")}
        print(json.dumps(line))
    return 0


def largest_single_gpu_n(pkg, m):
    """Largest round n whose (2m+6)-vector arena (+ slack) fits the free memory of ONE GPU."""
    free, total = pkg.mem_info()
    per_elem = 8.0 * (2 * m + 6)
    n = int(0.88 * (free - (6 << 30)) / per_elem)  # head-room: the comparator must never drive the GPU out of memory
    return max(1_000_000, (n // 10_000_000) * 10_000_000)


def run_config4(b, args):
    """BASELINE config 4: Rosenbrock n=2e9, m=20 over the N GPUs of the box; comparator = ONE GPU at the largest n that
    fits it with m=20, measured in this run by rank 0 (per-element throughput, no extrapolation from another size)."""
    pkg, rank, world = b.pkg, b.rank, b.world
    m = 20 if args.hist == M else args.hist
    n_global = 2_000_000_000 if args.size == N_GLOBAL else args.size
    K, W = min(args.steps, 10), 3
    need_gb = 8.0 * (2 * m + 6) * n_global / world / 1e9
    free, total = pkg.mem_info()
    if need_gb * 1e9 > free - (4 << 30):
        if rank == 0:
            print(json.dumps({"metric": "L-BFGS iterations/sec (FP64) at n=%.0e, m=%d" % (n_global, m), "value": None,
                              "n_gpus": world, "config": {"workload": "config 4"},
                              "unavailable": "n=%d, m=%d needs %.0f GB per GPU on %d GPUs; %.0f GB are free" %
                                             (n_global, m, need_gb, world, free / 1e9)}))
        return 0
    x0, n_local = b.shard_x0(n_global)
    run = b.measure(n_global, m, "compact", x0, K, W, sample_clocks=True, graph=1)
    two = b.measure(n_global, m, "two_loop", x0, K, W, graph=1, profile_kernels=False) if not args.single_variant else None
    x0.free()
    del x0
    pkg.trim_memory()
    comparator = None
    if rank == 0:
        try:
            n1 = largest_single_gpu_n(pkg, m)
            single = Bench.__new__(Bench)
            single.__dict__.update(b.__dict__)
            single.comm, single.dist, single.world = None, None, 1
            off0 = pkg.PinnedArray(n1)
            pkg.x0_uniform(n1, -2.0, 2.0, seed=42, offset=0, out=off0.array)
            r1 = single.measure(n1, m, "compact", off0, K, W, graph=1, profile_kernels=False)
            off0.free()
            comparator = {"n": n1, "n_gpus": 1, "value": r1["value"], "ms_per_step": r1["ms_per_step"],
                          "elements_per_second": r1["value"] * n1,
                          "what": "ONE GPU of this box, same code, m=%d, the largest round n that fits it" % m}
        except Exception as e:
            comparator = {"unavailable": repr(e)}
    b.barrier()
    if rank == 0:
        line = {"metric": "L-BFGS iterations/sec (FP64) at n=%.0e, m=%d" % (n_global, m), "value": run["value"], "unit": "iterations/s",
                "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": run["ms_per_step"], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "BASELINE config 4: Rosenbrock n=%d, m=%d, Wolfe, compact direction (m > 13: stand-alone pass A + combine kernels), CUDA-graph "
                                       "loop, contiguous shards x%d with 1-element halo + packed exchange" % (n_global, m, world),
                           "trials_per_step": run["trials_per_step"],
                           "cache": "each of the 2m+6 vectors is %.2f GB per GPU" % (8.0 * n_local / 1e9)},
                "gpu_launches": run["gpu_launches"], "clocks": run["clocks"], "roofline": run["roofline"],
                "two_loop": {k: two[k] for k in ("value", "ms_per_step", "whole_step")} if two else None,
                "single_gpu_comparator": comparator, "final": run["final"]}
        if comparator and "elements_per_second" in comparator:
            line["speedup_vs_one_gpu_per_element"] = run["value"] * n_global / comparator["elements_per_second"]
        print(json.dumps(line))
    return 0


def run_config5(b, args):
    """BASELINE config 5: history sweep m = 3..50 at n=1e8 over the N GPUs, two-loop vs compact, roofline fraction per m."""
    pkg, rank, world = b.pkg, b.rank, b.world
    n_global = args.size
    hists = [int(v) for v in args.hists.split(",")]
    K = min(args.steps, 15)
    x0, n_local = b.shard_x0(n_global)
    sweep = []
    for m in hists:
        for direction in ("two_loop", "compact"):
            r = b.measure(n_global, m, direction, x0, K, 3, graph=1, profile_kernels=(direction == "compact"))
            sweep.append({"m": m, "direction": direction, "value": r["value"], "ms_per_step": r["ms_per_step"],
                          "trials_per_step": r["trials_per_step"], "algorithmic_GB_per_step_per_gpu": r["whole_step"]["algorithmic_GB_per_step"],
                          "achieved_GBps_per_gpu": r["whole_step"]["achieved_GBps"], "frac_of_measured_peak": r["whole_step"]["frac_of_peak"],
                          "frac_of_8TBps": r["whole_step"]["frac_of_8TBps_nominal"], "f": r["final"]["f"],
                          "kernels": {k: {"avg_launch_ms": v["avg_launch_ms"], "frac_of_peak": v["frac_of_peak"]}
                                      for k, v in r.get("roofline", {}).get("kernels", {}).items()} or None})
        pkg.trim_memory()
    if rank == 0:
        best = max((s for s in sweep if s["m"] == 10 and s["direction"] == "compact"), key=lambda s: s["value"], default=sweep[0])
        print(json.dumps({"metric": "L-BFGS iterations/sec (FP64) at n=%.0e, history sweep" % n_global, "value": best["value"],
                          "unit": "iterations/s", "n_gpus": world, "steps": K, "warmup": 3, "ms_per_step": best["ms_per_step"],
                          "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": {"workload": "BASELINE config 5: Rosenbrock n=%d, Wolfe, m in %s, explicit two-loop vs compact "
                                                 "(fused flow for m <= 13), CUDA-graph loop, x%d GPUs; value = the m=10 compact entry" %
                                                 (n_global, hists, world)},
                          "peak": b.peak, "peak_source": b.peak_src, "sweep": sweep}))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=12)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", type=int, default=2, choices=(2, 4, 5))
    ap.add_argument("--size", type=int, default=N_GLOBAL, help=argparse.SUPPRESS)  # problem size n
    ap.add_argument("--hist", type=int, default=M, help=argparse.SUPPRESS)         # history size m
    ap.add_argument("--hists", default="3,5,10,20,30,50", help=argparse.SUPPRESS)  # --config 5
    ap.add_argument("--sustain", type=float, default=3.0, help=argparse.SUPPRESS)  # seconds of the sustained leg (0 = off)
    ap.add_argument("--no-cpu-baseline", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--graph", type=int, default=1, help=argparse.SUPPRESS)
    ap.add_argument("--direction", default="compact", help=argparse.SUPPRESS)
    ap.add_argument("--single-variant", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    b = Bench(args)
    try:
        if args.config == 4:
            return run_config4(b, args)
        if args.config == 5:
            return run_config5(b, args)
        return run_config2(b, args)
    finally:
        b.close()


if __name__ == "__main__":
    sys.exit(main())
