#!/bin/bash
# DRAM bytes per launch of the decoder GEMMs (video-prefix run, M = 25 728) with / without the K-sliced down_proj and the
# N-sliced gate|up:   tools/ab_gemm_split.sh "<ksplit> <nsplit>" ...      e.g.  "1 1" "2 1" "2 6"
for cfg in "$@"; do set -- $cfg; export BLIM_GEMM_KSPLIT=$1 BLIM_GEMM_NSPLIT=$2
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "regex:gemm_tcgen05|rmsnorm" --launch-skip ${SKIP:-166} --launch-count ${COUNT:-12} --csv --log-file gpurun_out/split_$1_$2.csv python bench.py --n 96 --warmup 0 --steps 1 --no-e2e --no-cpu-baseline --no-parity > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader([l for l in open('gpurun_out/split_$1_$2.csv') if l.startswith('"')]))
h=rows[0]
out={}
for r in rows[1:]:
    out.setdefault((int(r[h.index('ID')]), r[h.index('Kernel Name')][:60]),{})[r[h.index('Metric Name')][:16]]=float(r[h.index('Metric Value')].replace(',',''))
print('ksplit=$1 nsplit=$2')
for (i,k),m in sorted(out.items()):
    print(i, k, 'ms %.3f' % (m.get('gpu__time_durati',0)/1e6), 'read MB %.0f' % (m.get('dram__bytes_read',0)/1e6), 'write MB %.0f' % (m.get('dram__bytes_writ',0)/1e6))
PY
done
