#!/bin/bash
# Tuning sweep of the fused kernels' tile configuration (LBFGSB200_AG_TILE / LBFGSB200_CT_TILE = "T,NS", LBFGSB200_CT_HALO)
# on one GPU: n = 1e8, Wolfe, per history size.  One JSON summary line per configuration.
cd "$(dirname "$0")/.."
run() { # m, extra env...
  local m=$1; shift
  env "$@" python bench.py --hist $m --steps 12 --warmup 3 --no-cpu-baseline --single-variant --sustain 0 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read()); k=r['roofline']['kernels']
print(json.dumps({'m':$m,'env':'$*','value':round(r['value'],2),'trials':r['config']['trials_per_step'],'ms':round(r['ms_per_step'],3),'kernels':{a:[round(b['avg_launch_ms'],3),round(b['frac_of_peak'],3)] for a,b in k.items()}}))"
}
for cfg in "$@"; do
  IFS=: read m envs <<< "$cfg"
  run $m ${envs//;/ }
done
