for cfg in "0 40" "1 40" "0 64" "0 24"; do set -- $cfg; export BLIM_GEMM_HINTS=$1 BLIM_GEMM_SB_MB=$2
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "regex:gemm_tcgen05" --launch-skip 124 --launch-count 4 --csv --log-file gpurun_out/hints_$1_$2.csv python bench.py --n 96 --warmup 0 --steps 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader([l for l in open('gpurun_out/hints_$1_$2.csv') if l.startswith('"')]))
h=rows[0]
out={}
for r in rows[1:]:
    out.setdefault((r[h.index('ID')], r[h.index('Kernel Name')][26:60]),{})[r[h.index('Metric Name')]]=r[h.index('Metric Value')]+r[h.index('Metric Unit')]
print('hints=$1 sb=$2', out)
PY
done
