#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_parity_gpu.py tests/test_benchmarked_configs_gpu.py -x -q > gpurun_out/c_tests.log 2>&1; echo "pytest rc $?" >> gpurun_out/c_tests.log
tail -3 gpurun_out/c_tests.log
run() { tag=$1; shift; env "$@" python bench.py --no-cpu-baseline --no-gpu-eager-baseline --no-optimizer > gpurun_out/c_bench_$tag.json 2> gpurun_out/c_bench_$tag.err; python -c "
import json;d=json.loads(open('gpurun_out/c_bench_$tag.json').read().strip().splitlines()[-1]);print('$tag',d['ms_per_step'],d['value'],d['loss'])"; }
run default A=1
run lag4 OF_SIDE_LAG=4
run lag1 OF_SIDE_LAG=1
run noside OF_WGRAD_SIDE=0
bash tools/ncu_step_hbm.sh gpurun_out/r02_launches_hbm.csv
