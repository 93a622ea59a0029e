#!/bin/bash
# Launch list (gpu__time_duration per kernel) of exactly ONE CUDA-graph replay of the training micro-step, for profiles/.
# usage: tools/ncu_step_list.sh <out.csv>     (bench.py brackets one replay with cudaProfilerStart/Stop under OF_PROFILE_STEP=1)
set -e
OUT=${1:-gpurun_out/launches.csv}
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1
OF_PROFILE_STEP=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$OUT" \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
