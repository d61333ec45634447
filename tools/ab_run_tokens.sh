for rt in 0 40960 49152 65536 0; do
python bench.py --n 125 --run-tokens $rt --steps 3 --warmup 2 --no-e2e --no-parity --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
r=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('n125 run_tokens $rt', 'ms/step', round(r['ms_per_step'],1), r['clocks']['sm_mhz'], 'launches', r['gpu_launches'], 'kernels', round(sum(x['ms_per_step'] for x in r['roofline']['by_kernel'].values()),1))"
done
for rt in 0 49152 65536 0 49152; do
python bench.py --run-tokens $rt --steps 2 --warmup 2 --no-e2e --no-parity --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
r=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('full run_tokens $rt', 'pairs/s', round(r['value'],1), r['clocks']['sm_mhz'], 'launches', r['gpu_launches'], 'frac', round(r['roofline']['frac'],3))"
done
