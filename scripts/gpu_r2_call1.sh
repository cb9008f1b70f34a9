#!/bin/bash
# Round-2 GPU call 1 (1 GPU): smoke of the fused flow, the whole -m gpu suite, large-size parity, A/B bench, CUDA goldens.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/c1_gpus.txt 2>&1
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke > $O/c1_smoke.log 2>&1; echo "smoke rc=$?"; tail -8 $O/c1_smoke.log
echo "== pytest (all but large)"; timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_large.py > $O/c1_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $O/c1_pytest.log
echo "== large parity"; rm -f $O/large_parity.jsonl; timeout 1200 python -m pytest tests/test_gpu_large.py -q > $O/c1_large.log 2>&1; echo "large rc=$?"; tail -12 $O/c1_large.log
echo "== bench fused"; timeout 400 python bench.py --no-cpu-baseline > $O/c1_bench_fused.json 2> $O/c1_bench_fused.err; echo "rc=$?"; cut -c1-700 $O/c1_bench_fused.json; tail -3 $O/c1_bench_fused.err
echo "== bench unfused"; LBFGSB200_FUSED=0 timeout 400 python bench.py --no-cpu-baseline --single-variant > $O/c1_bench_unfused.json 2> $O/c1_bench_unfused.err; echo "rc=$?"; cut -c1-400 $O/c1_bench_unfused.json
echo "== cuda goldens"; timeout 600 python oracle/make_golden_cuda.py $O/cuda_reference_traces.json > $O/c1_cudagolden.log 2>&1; echo "rc=$?"; tail -8 $O/c1_cudagolden.log
