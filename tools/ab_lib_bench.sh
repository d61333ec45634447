#!/bin/bash
# same-box interleaved A/B of builds / environment settings on the default C2 job:
#   tools/ab_lib_bench.sh "NAME ENV=VAL ..." ...      (BLIM_LIB=<path> selects another build of the shared object)
for cfg in "$@"; do
  set -- $cfg; name=$1; shift
  env "$@" X_=1 python bench.py --steps 2 --warmup 2 --no-e2e --no-parity --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
r=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$name', round(r['value'],1), r['clocks']['sm_mhz'], {k:(round(x['ms_per_step']),x['tflops'] and round(x['tflops'])) for k,x in r['roofline']['by_kernel'].items() if k in ('qkv_rope','o_proj','down_proj','gate_up_swiglu')})"
done
