#!/bin/bash
# The GPU-side checks used during round 1, as gpurun one-liners (run from the repo root in the build container).
# Usage: scripts/gpu_checks.sh tests | bench | ncu | multi2 | multi8 | cudagolden | timeline
set -e
GPURUN=${GPURUN:-/usr/local/graft/bin/gpurun}
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
case "$1" in
tests)  $GPURUN --timeout 900 -- 'timeout 700 python -m pytest tests -m gpu -q 2>&1 | tail -5; timeout 120 python __graft_entry__.py smoke | tail -3' ;;
bench)  $GPURUN --timeout 900 -- 'timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err; timeout 200 python bench.py --impl reference > gpurun_out/bench_ref.json; cat gpurun_out/bench.json | cut -c1-600' ;;
ncu)    # ncu cannot profile kernel nodes of graphs with conditional nodes: capture the host-stepped loop (--graph 0)
        $GPURUN --timeout 1500 -- 'B="python bench.py --steps 4 --warmup 12 --no-cpu-baseline --single-variant --graph 0"; $B > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 215 -c 80 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu1.log 2>&1; ncu --set full --clock-control none --import-source on -k regex:"k_gram|k_combine|k_trial|k_accept" -s 62 -c 10 -o gpurun_out/prof $B > gpurun_out/ncu2.log 2>&1; ls -la gpurun_out' ;;
multi2) $GPURUN --gpus 2 --timeout 900 -- "timeout 500 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -2; timeout 200 $T --nproc-per-node 2 --master-port 29571 bench.py --gpus 2 --no-cpu-baseline 2>/dev/null | grep '^{' | cut -c1-400" ;;
multi8) $GPURUN --gpus 8 --timeout 1500 -- "for N in 2 4 8; do timeout 300 $T --nproc-per-node \$N --master-port 2953\$N bench.py --gpus \$N --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/bench_n\$N.json; done; timeout 600 $T --nproc-per-node 8 --master-port 29540 bench.py --gpus 8 --size 2000000000 --hist 20 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/config4_n8.json; timeout 400 $T --nproc-per-node 8 --master-port 29555 benchmarks/msweep_multi.py 2>/dev/null | grep '^{' > gpurun_out/msweep_8gpu.jsonl; ls -la gpurun_out" ;;
cudagolden) # the reference's own CUDA solvers (oracle/_ref/libref_cuda_*.so, built here by `make -C oracle cudaref`) on a B200
        $GPURUN --timeout 600 -- 'timeout 400 python oracle/make_golden_cuda.py gpurun_out/cuda_reference_traces.json | tail -12'
        echo "then: cp gpurun_out/cuda_reference_traces.json tests/golden/" ;;
timeline) $GPURUN --gpus 2 --timeout 600 -- "timeout 120 python benchmarks/timeline.py 2>/dev/null | grep '^{' > gpurun_out/timeline.jsonl; timeout 120 $T --nproc-per-node 2 --master-port 29601 benchmarks/timeline.py --gpus 2 --size 25000000 2>/dev/null | grep '^{' >> gpurun_out/timeline.jsonl; cat gpurun_out/timeline.jsonl | cut -c1-300" ;;
*) echo "usage: $0 tests|bench|ncu|multi2|multi8|cudagolden|timeline"; exit 2 ;;
esac
