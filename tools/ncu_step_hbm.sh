#!/bin/bash
# Launch list of ONE graph replay with DRAM bytes per launch (for tools/hbm_kernels.py).  usage: tools/ncu_step_hbm.sh <out.csv>
set -e
OUT=${1:-gpurun_out/launches_hbm.csv}
OF_PROFILE_STEP=1 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file "$OUT" \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline --no-optimizer > gpurun_out/ncu_hbm_bench.log 2>&1
