#!/bin/bash
# Retry a gpurun call while the pod answers "transient" / busy (nothing is charged for those): scripts/gpurun_retry.sh <gpurun args...>
out=gpurun_out/retry.log
mkdir -p gpurun_out
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > $out 2>&1
  if ! grep -q "status=transient\|status=busy\|rc=3\b" $out; then break; fi
  sleep 90
done
tail -60 $out
