#!/bin/bash
# quick A/B of a host-side change: whole-model parity tests + the train bench with and without the flag given in $1 (e.g. OF_WGRAD_SIDE)
mkdir -p gpurun_out
FLAG=${1:-OF_WGRAD_SIDE}
timeout 900 python -m pytest tests/test_model_parity_gpu.py tests/test_benchmarked_configs_gpu.py -x -q > gpurun_out/b_tests.log 2>&1; echo "pytest rc $?" >> gpurun_out/b_tests.log
tail -3 gpurun_out/b_tests.log
python bench.py --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/b_bench_on.json 2> gpurun_out/b_bench_on.err; python -c "
import json;d=json.loads(open('gpurun_out/b_bench_on.json').read().strip().splitlines()[-1]);print('ON ',d['ms_per_step'],d['value'],d['loss'])"
env $FLAG=0 python bench.py --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/b_bench_off.json 2> gpurun_out/b_bench_off.err; python -c "
import json;d=json.loads(open('gpurun_out/b_bench_off.json').read().strip().splitlines()[-1]);print('OFF',d['ms_per_step'],d['value'],d['loss'])"
tail -5 gpurun_out/b_bench_on.err
