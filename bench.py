#!/usr/bin/env python
"""bench.py -- L-BFGS iterations/s (FP64) at n=1e8, m=10, Wolfe line search (BASELINE.json config 2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one L-BFGS iteration (two-loop recursion over the full history, a Wolfe line search
with fused trial evaluations, the accept/update pass) of the Rosenbrock objective, x0 ~ U(-2,2)
drawn like the reference mains (mt19937(42)).  Inputs are synthetic; every vector is 0.8 GB, far
larger than the 126 MB L2, so nothing survives in cache between passes.

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline     -- the dominant streaming kernel of the step (plus a table of all of them): achieved
                  GB/s = algorithmic bytes per launch / CUDA-event time of every launch in a separate
                  instrumented run of the same iterations, against MEASURED_PEAKS.json:hbm_gbs
  variants     -- the same measurement with the other direction algorithm (explicit two-loop
                  recursion vs the compact/Gram form that reads the history once)
  cpu_baseline -- the UNMODIFIED reference (oracle/_ref, sequential outer loop + the CUDA tree's
                  Wolfe search = the hybrid oracle) on one host core, on a bounded sample
  e2e          -- same metric through the host-buffer API: create + H2D x0 + W+K iterations +
                  D2H x, wall clock
  reference_cuda_on_this_gpu -- the reference's own CUDA solver for this configuration
                  (parallel-implementation/L-BFGS-Wolfe.cu, unmodified, cross-compiled for sm_100 under
                  oracle/_ref) timed on the same GPU on the same bounded sample as cpu_baseline

--impl reference times only the reference's CPU implementation (rank 0; other ranks exit 0).
"""
import argparse
import importlib.util
import json
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
CPU_SAMPLE_N = 1_000_000  # bounded CPU sample: 1% of the workload, scaled linearly in n


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
            time.sleep(0.1)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples if len(s) >= 7 for i in range(4)
                          if s[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def cpu_reference_rate(steps, warmup, n_sample=CPU_SAMPLE_N):
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
            "sample": "n=%d (1/%d of the workload) x %d steady-state iterations after %d warm-up, measured "
                      "%.3f s/iteration, scaled linearly in n to n=%d" % (n_sample, N_GLOBAL // n_sample, steps, warm,
                                                                        per_iter, N_GLOBAL),
            "host_cores_available": os.cpu_count(), "seconds_per_iteration_at_sample": per_iter}


def cuda_reference_rate(steps, warmup, n_sample=CPU_SAMPLE_N):
    """The reference's own CUDA solver for this configuration (parallel-implementation/L-BFGS-Wolfe.cu, unmodified,
    cross-compiled for sm_100: oracle/_ref/libref_cuda_wolfe.so) on the SAME B200: steady-state seconds per
    iteration on n_sample elements, scaled linearly to N_GLOBAL.  Its iteration ships x and the gradient across
    PCIe and evaluates f / grad on one host core, so linear scaling in n is exact to first order.
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
    t_warm = run(warm)
    t_all = run(warm + steps)
    per_iter = max(t_all - t_warm, 1e-9) / steps
    return {"value": (1.0 / per_iter) * n_sample / N_GLOBAL, "unit": "iterations/s", "kind": "reference (CUDA tree)",
            "source": "parallel-implementation/L-BFGS-Wolfe.cu, unmodified, nvcc defaults, sm_100, cuBLAS; 1 GPU + 1 host core",
            "sample": "n=%d (1/%d of the workload) x %d steady-state iterations after %d warm-up, measured %.4f "
                      "s/iteration, scaled linearly in n to n=%d" % (n_sample, N_GLOBAL // n_sample, steps, warm, per_iter,
                                                                     N_GLOBAL),
            "seconds_per_iteration_at_sample": per_iter}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps = min(args.steps, 30)
    cb = cpu_reference_rate(steps, min(args.warmup, 12))
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
            "unit": "iterations/s", "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 12),
            "ms_per_step": 1e3 / cb["value"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "Rosenbrock n=1e8, m=10, Wolfe line search (hybrid CPU reference: "
                                   "sequential-implementation/lbfgs.cpp + parallel-implementation/line_search.cpp), "
                                   "bounded sample scaled linearly in n"},
            "cpu_baseline": cb, "config1": config1,
            "e2e": {"value": cb["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def main():
    global M
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=12)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--size", type=int, default=N_GLOBAL, help=argparse.SUPPRESS)  # problem size n
    ap.add_argument("--hist", type=int, default=M, help=argparse.SUPPRESS)         # history size m
    ap.add_argument("--no-cpu-baseline", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--graph", type=int, default=1, help=argparse.SUPPRESS)
    ap.add_argument("--direction", default="compact", help=argparse.SUPPRESS)
    ap.add_argument("--single-variant", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_global = args.size
    M = args.hist
    K, W = args.steps, max(args.warmup, 3)
    FILL = max(0, M + 2 - W)  # extra untimed iterations before the warm-up so the history is full (h = m)

    pkg = load_pkg()
    L = pkg.lib()
    if L.lbfgsb200_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback")
    pkg._check(L.lbfgsb200_set_device(local_rank), "set_device")

    dist = None
    comm = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        ids = [pkg.Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        comm = pkg.Comm(ids[0], rank, world)

    def barrier():
        if dist is not None:
            dist.barrier()
        L.lbfgsb200_device_sync()

    def max_over_ranks(v):
        if dist is None:
            return v
        import torch
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    off, n_local = pkg.shard_range(n_global, rank, world)
    x0_pinned = pkg.PinnedArray(n_local)
    out_pinned = pkg.PinnedArray(n_local)
    pkg.x0_uniform(n_local, -2.0, 2.0, seed=42, offset=off, out=x0_pinned.array)
    barrier()  # ranks generate shards of different offsets (mt19937 skip-ahead): line up before any exchange

    peak, peak_src = measured_peak()
    V = 8.0 * n_local
    h = M
    # algorithmic bytes per launch of each streaming-kernel class
    fused = os.environ.get("LBFGSB200_FUSED", "1") != "0"
    # fused compact flow (default): k_accept_gram reads the 2(h-1) kept history rows + x, d, g and writes x, g, s, y;
    # k_combine_trial reads the 2h+1 basis vectors + x and writes d (the first line-search trial rides on it)
    class_bytes = {"two_loop_pass": (8 * h - 1) * V / (2 * h), "gram_rows": ((2 * (h - 1) + 7) if fused else (2 * h + 1)) * V,
                   "combine": ((2 * h + 3) if fused else (2 * h + 2)) * V, "trial": 2 * V, "accept": 7 * V}
    class_kernel = {"two_loop_pass": "k_two_loop_pass", "gram_rows": "k_accept_gram" if fused else "k_gram_tma2d",
                    "combine": "k_combine_trial" if fused else "k_combine", "trial": "k_trial", "accept": "k_accept"}
    traffic_db = {}
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic_db = json.load(open(tpath))
        except Exception:
            traffic_db = {}

    def measure(direction, sample_clocks):
        """Device-resident timed region of K iterations after W warm-up iterations, then a separate
        instrumented run of K more iterations with a CUDA-event pair around every streaming kernel."""
        prm = pkg.default_params(FLAVOR, line_search=LINE_SEARCH, m=M, max_iterations=10 ** 9, tolerance=0.0,
                                 use_graph=args.graph, direction=direction)
        solver = pkg.Solver(OBJECTIVE, n_global, prm, comm=comm, trace_rows=FILL + W + 2 * K + 8)
        solver.set_x0(x0_pinned.array)
        if FILL:
            solver.iterate(FILL)  # fill the (s, y) history so every timed step uses all m pairs
        solver.iterate(W)  # warm-up
        launches0 = solver.result()["kernel_launches"]
        sampler = ClockSampler(local_rank) if (rank == 0 and sample_clocks) else None
        if sampler:
            sampler.start()
        barrier()
        t0 = time.perf_counter()
        solver.iterate(K)
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        if sampler:
            sampler.stop_flag = True
        res = solver.result()
        dev_ms = max_over_ranks(res["device_ms"])
        wall_ms = max_over_ranks(wall_ms)
        launches = res["kernel_launches"] - launches0
        rows = solver.trace()[FILL + W:FILL + W + K]
        trials = float(np.sum(rows[:, 4])) if len(rows) else 0.0
        bytes_step = res["bytes_moved"] / K  # local shard, algorithmic
        _, classes = solver.iterate_profiled(K)
        res_p = solver.result()
        solver.destroy()
        kernels = {}
        for name, c in classes.items():
            if name in class_bytes and c["launches"] > 0 and c["ms"] > 0:
                ms = c["ms"] / c["launches"]
                gbs = class_bytes[name] / (ms * 1e-3) / 1e9
                kernels[class_kernel[name]] = {"launches_per_step": c["launches"] / K, "avg_launch_ms": ms,
                                               "algorithmic_bytes_per_launch": class_bytes[name], "achieved_GBps": gbs,
                                               "frac_of_peak": gbs / peak, "ms_per_step": c["ms"] / K,
                                               "traffic": (traffic_db[class_kernel[name]]["dram_over_algorithmic"] * class_bytes[name]
                                                           if "dram_over_algorithmic" in traffic_db.get(class_kernel[name], {}) else None)}
        dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
        step_gbs = bytes_step / (dev_ms / K * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_GBps"], "peak": peak, "unit": "GB/s",
                    "frac": kernels[dom]["frac_of_peak"], "traffic": kernels[dom]["traffic"], "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": kernels[dom]["algorithmic_bytes_per_launch"],
                    "avg_launch_ms": kernels[dom]["avg_launch_ms"], "kernels": kernels,
                    "whole_step": {"algorithmic_GB_per_step": bytes_step / 1e9, "achieved_GBps": step_gbs,
                                   "frac_of_peak": step_gbs / peak, "frac_of_8TBps_nominal": step_gbs / 8000.0},
                    "instrumented_ms_per_step": res_p["device_ms"] / K}
        return {"value": K / (dev_ms / 1e3), "ms_per_step": dev_ms / K, "wall_ms_per_step": wall_ms / K,
                "gpu_launches": int(launches), "trials_per_step": trials / K if K else None, "roofline": roofline,
                "clocks": sampler.summary() if sampler else None,
                "final": {"f": res_p["f"], "gnorm": res_p["gnorm"], "iterations": res_p["iterations"]}}

    main_run = measure(args.direction, True)
    other = "two_loop" if args.direction == "compact" else "compact"
    other_run = measure(other, False) if not args.single_variant else None
    params = pkg.default_params(FLAVOR, line_search=LINE_SEARCH, m=M, max_iterations=10 ** 9, tolerance=0.0,
                                use_graph=args.graph, direction=args.direction)

    # ---------------- end to end through the host-buffer API ----------------
    barrier()
    t0 = time.perf_counter()
    s2 = pkg.Solver(OBJECTIVE, n_global, params, comm=comm, trace_rows=0)
    s2.set_x0(x0_pinned.array)      # H2D of x0 from pinned host memory
    s2.iterate(FILL + W + K)
    s2.x(out=out_pinned.array)      # D2H of the result
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    s2.destroy()
    e2e = {"value": (FILL + W + K) / e2e_s, "unit": "iterations/s", "h2d_bytes_per_step": 8.0 * n_global / (FILL + W + K),
           "d2h_bytes_per_step": 8.0 * n_global / (FILL + W + K),
           "what": "create + H2D x0 (pinned) + %d iterations from a cold history + D2H x; wall clock, max over ranks; "
                   "byte counts are the one-off 8n-byte copies amortised over the iterations of the call" % (FILL + W + K)}

    cpu_baseline = None
    cuda_reference = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = cpu_reference_rate(min(K, 20), 12)
        try:
            cuda_reference = cuda_reference_rate(min(K, 20), 12)
        except Exception as e:  # the comparison is informative, never fatal to the bench line
            cuda_reference = {"unavailable": repr(e)}

    if rank == 0:
        line = {"metric": "L-BFGS iterations/sec (FP64) at n=%.0e, m=%d" % (n_global, M) if (n_global != N_GLOBAL or M != 10) else "L-BFGS iterations/sec (FP64) at n=1e8, m=10", "value": main_run["value"], "unit": "iterations/s",
                "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": main_run["ms_per_step"], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "Rosenbrock n=%d, m=%d, Wolfe line search (C2=0.7, safeguarded cubic), "
                                       "x0~U(-2,2) mt19937(42), %s direction, %s" %
                                       (n_global, M, args.direction, "CUDA-graph loop" if args.graph else "host-stepped loop"),
                           "parallelism": "contiguous shards x%d, 1-element halo + packed all-gather per sync" % world
                                          if world > 1 else "single GPU",
                           "cache": "inputs larger than L2 (each of the 2m+6 vectors is %.2f GB per GPU)" % (V / 1e9),
                           "trials_per_step": main_run["trials_per_step"],
                           "history_fill_iterations_before_warmup": FILL},
                "wall_ms_per_step": main_run["wall_ms_per_step"], "gpu_launches": main_run["gpu_launches"],
                "clocks": main_run["clocks"], "roofline": main_run["roofline"], "e2e": e2e, "cpu_baseline": cpu_baseline, "reference_cuda_on_this_gpu": cuda_reference,
                "final": main_run["final"]}
        if other_run is not None:
            line["variants"] = {other: {k: other_run[k] for k in ("value", "ms_per_step", "gpu_launches", "trials_per_step", "final")}}
            line["variants"][other]["roofline"] = {k: other_run["roofline"][k] for k in ("kernel", "achieved", "frac", "whole_step")}
        print(json.dumps(line))
    if comm is not None:
        comm.destroy()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
