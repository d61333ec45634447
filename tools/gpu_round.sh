#!/bin/bash
# One GPU-box visit: parity tests, the default bench (+ reference arm), the ncu launch list of the bench command and one
# `ncu --set full` capture of a whole decoder layer of full-size launches.  Usage: tools/gpu_round.sh <tag> [quick]
TAG=${1:-vX}
MODE=${2:-all}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t_all_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_all_$TAG.log
( time python bench.py ) > gpurun_out/bench_default_$TAG.log 2> gpurun_out/bench_default_err_$TAG.log; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_default_$TAG.log
python bench.py --impl reference > gpurun_out/bench_reference_$TAG.log 2>/dev/null
if [ "$MODE" = "all" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_c2_n96_$TAG.csv \
      python bench.py --n 96 --warmup 0 --steps 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_bench_$TAG.log 2>&1
  ncu --set full --clock-control none --import-source on -k 'regex:gemm_tcgen05|attention_tc|rmsnorm' --launch-skip 230 --launch-count 14 \
      -f -o gpurun_out/prof_layer_$TAG python bench.py --n 96 --warmup 0 --steps 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_layer_$TAG.log 2>&1
  ls -la gpurun_out/*$TAG*
fi
