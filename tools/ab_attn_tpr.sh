#!/bin/bash
# tests + same-box A/B of one vs two softmax threads per row in the warp-specialised attention kernel (BLIM_ATTN=ws2)
BLIM_ATTN=ws2 timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_edge_gpu.py tests/test_shapes_gpu.py tests/test_pipeline_gpu.py -m gpu -q -x 2>&1 | tail -3
for v in ws ws2 ws ws2; do
  if [ $v = ws ]; then unset BLIM_ATTN; else export BLIM_ATTN=$v; fi
  python bench.py --steps 2 --warmup 2 --no-e2e --no-parity --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
r=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('attn $v', round(r['value'],1), r['clocks']['sm_mhz'], 'attention ms/step', round(r['roofline']['by_kernel']['attention']['ms_per_step'],1))"
done
