#!/bin/bash
# One rank's share of an 8-GPU C2 job on ONE GPU (bench.py --n 125: 125 videos / texts = the prefix owners of one rank):
# same-box A/B of settings that only matter for short decoder runs.   tools/ab_small_bench.sh "ENV=VAL ..." "ENV=VAL ..." ...
for cfg in "$@"; do
  env $cfg python bench.py --n 125 --steps 3 --warmup 2 --no-e2e --no-parity --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
r=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$cfg', 'ms/step', round(r['ms_per_step'],1), 'pairs/s', round(r['value'],1), r['clocks']['sm_mhz'], 'launches', r['gpu_launches'], {k:round(x['ms_per_step'],1) for k,x in r['roofline']['by_kernel'].items()}, 'kernels sum', round(sum(x['ms_per_step'] for x in r['roofline']['by_kernel'].values()),1))"
done
