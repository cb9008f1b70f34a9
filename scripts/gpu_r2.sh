#!/bin/bash
# Round-2 GPU stages (run under gpurun): scripts/gpu_r2.sh <stage> [<stage> ...]
#   smoke | tests | large | bench | bench_unfused | ab | cudagolden | multi | ncu
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
O=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for stage in "$@"; do
  echo "== $stage"
  case "$stage" in
    smoke) timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "rc=$?"; tail -6 $O/smoke.log ;;
    tests) timeout 2400 python -m pytest tests -m gpu -q --deselect tests/test_gpu_large.py > $O/pytest.log 2>&1; echo "rc=$?"; tail -25 $O/pytest.log ;;
    testsx) timeout 2400 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_large.py > $O/pytest.log 2>&1; echo "rc=$?"; tail -40 $O/pytest.log ;;
    large) rm -f $O/large_parity.jsonl; timeout 1500 python -m pytest tests/test_gpu_large.py -q > $O/large.log 2>&1; echo "rc=$?"; tail -12 $O/large.log ;;
    bench) timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; echo "rc=$?"; cut -c1-300 $O/bench.json; tail -3 $O/bench.err ;;
    benchq) timeout 400 python bench.py --no-cpu-baseline --single-variant --sustain 0 > $O/benchq.json 2> $O/benchq.err; echo "rc=$?"; cut -c1-300 $O/benchq.json; tail -3 $O/benchq.err ;;
    bench_unfused) LBFGSB200_FUSED=0 timeout 400 python bench.py --no-cpu-baseline --single-variant --sustain 0 > $O/bench_unfused.json 2> $O/bench_unfused.err; echo "rc=$?"; cut -c1-300 $O/bench_unfused.json ;;
    bench_ref) timeout 900 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err; echo "rc=$?"; cut -c1-600 $O/bench_ref.json ;;
    cudagolden) timeout 600 python oracle/make_golden_cuda.py $O/cuda_reference_traces.json > $O/cudagolden.log 2>&1; echo "rc=$?"; tail -8 $O/cudagolden.log ;;
    multi) timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_compat.py -q > $O/pytest_multi.log 2>&1; echo "rc=$?"; tail -25 $O/pytest_multi.log ;;
    scale) for N in 2 4 8; do [ $N -le $(nvidia-smi -L | wc -l) ] && timeout 400 $T --nproc-per-node $N --master-port 2953$N bench.py --gpus $N --no-cpu-baseline --single-variant 2>$O/bench_n$N.err | grep '^{' > $O/bench_n$N.json; cut -c1-200 $O/bench_n$N.json; done ;;
    config4) N=$(nvidia-smi -L | wc -l); timeout 900 $T --nproc-per-node $N --master-port 29540 bench.py --gpus $N --config 4 2>$O/config4.err | grep '^{' > $O/config4_n$N.json; cut -c1-300 $O/config4_n$N.json; tail -3 $O/config4.err ;;
    config5) N=$(nvidia-smi -L | wc -l); timeout 900 $T --nproc-per-node $N --master-port 29555 bench.py --gpus $N --config 5 2>$O/config5.err | grep '^{' > $O/config5_n$N.json; cut -c1-300 $O/config5_n$N.json; tail -3 $O/config5.err ;;
    ncu) B="python bench.py --steps 12 --warmup 12 --no-cpu-baseline --single-variant --sustain 0 --graph 0"
         $B > $O/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 120 --csv --log-file $O/launches.csv $B > $O/ncu1.log 2>&1
         ncu --set full --clock-control none --import-source on -k regex:"k_accept_gram|k_combine_trial|k_trial" -s 60 -c 9 -o $O/prof $B > $O/ncu2.log 2>&1; ls -la $O | tail -5 ;;
    e2e) timeout 600 python benchmarks/e2e_breakdown.py 100000000 42 > $O/e2e_1e8.jsonl 2>&1; cat $O/e2e_1e8.jsonl | cut -c1-420
         timeout 600 python benchmarks/e2e_breakdown.py 12500000 42 > $O/e2e_125e5.jsonl 2>&1; cat $O/e2e_125e5.jsonl | cut -c1-420 ;;
    timeline) N=$(nvidia-smi -L | wc -l); timeout 120 python benchmarks/timeline.py --size 12500000 2>/dev/null | grep '^{' > $O/timeline.jsonl
         timeout 120 python benchmarks/timeline.py 2>/dev/null | grep '^{' >> $O/timeline.jsonl
         [ $N -ge 2 ] && timeout 200 $T --nproc-per-node $N --master-port 29601 benchmarks/timeline.py --gpus $N --size $((12500000*N)) 2>/dev/null | grep '^{' >> $O/timeline.jsonl
         cat $O/timeline.jsonl | cut -c1-900 ;;
    multi8) timeout 900 python -m pytest tests/test_gpu_multi.py -q -k "every_device or one_process" > $O/pytest_multi8.log 2>&1; echo "rc=$?"; tail -8 $O/pytest_multi8.log ;;
    ncu20) B="python bench.py --hist 20 --steps 4 --warmup 3 --no-cpu-baseline --single-variant --sustain 0 --graph 0"
         ncu --set full --clock-control none -k regex:"k_accept_gram|k_combine_trial" -s 50 -c 4 -o $O/prof20 $B > $O/ncu20.log 2>&1; ls -la $O/prof20* ;;
    *) echo "unknown stage $stage" ;;
  esac
done
