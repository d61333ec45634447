for cfg in "1 1" "1 6" "3 6" "1 8" "3 1" "1 1" "1 6"; do set -- $cfg; BLIM_GEMM_KSPLIT=$1 BLIM_GEMM_NSPLIT=$2 python bench.py --steps 2 --warmup 2 --no-e2e --no-parity --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
r=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('ksplit $1 nsplit $2', round(r['value'],1), r['clocks']['sm_mhz'], {k:(round(x['ms_per_step']),x['tflops'] and round(x['tflops'])) for k,x in r['roofline']['by_kernel'].items() if k in ('gate_up_swiglu','down_proj','o_proj')})"; done
