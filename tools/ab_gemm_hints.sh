#!/bin/bash
# DRAM bytes per launch of the decoder GEMMs (M = 25 728) under the L2 knobs of the tcgen05 GEMM:
#   tools/ab_gemm_hints.sh "<hints> <slab MB> <min m-tiles>" ...      e.g.  "0 40 2" "1 40 2" "0 4 1"
for cfg in "$@"; do set -- $cfg; export BLIM_GEMM_HINTS=$1 BLIM_GEMM_SB_MB=$2 BLIM_GEMM_SB_MIN=$3
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "regex:gemm_tcgen05" --launch-skip 124 --launch-count 4 --csv --log-file gpurun_out/hints_$1_$2_$3.csv python bench.py --n 96 --warmup 0 --steps 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader([l for l in open('gpurun_out/hints_$1_$2_$3.csv') if l.startswith('"')]))
h=rows[0]
out={}
for r in rows[1:]:
    out.setdefault((r[h.index('ID')], r[h.index('Kernel Name')][26:44]),{})[r[h.index('Metric Name')][:14]]=round(float(r[h.index('Metric Value')].replace(',',''))/1e6,1)
print('hints=$1 slab=$2MB min=$3', out)
PY
done
