#!/bin/bash
# same-box interleaved A/B of the scoring-path attention kernels: tools/ab_attn_bench.sh [workload]
WL=${1:-c2}
for v in tc2p ws tc2p ws; do
  if [ $v = ws ]; then unset BLIM_ATTN; else export BLIM_ATTN=$v; fi
  python bench.py --workload $WL --steps 2 --warmup 2 --no-e2e --no-parity --no-cpu-baseline ${EXTRA} 2>/dev/null | python -c "
import json,sys
r=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('attn $v', round(r['value'],1), r['clocks']['sm_mhz'], 'attention ms/step', round(r['roofline']['by_kernel']['attention']['ms_per_step'],1), 'share', round(r['roofline']['attention_share_of_step'],4), 'gemm frac', round(r['roofline']['frac'],3))"
done
