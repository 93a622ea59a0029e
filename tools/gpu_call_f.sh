#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lora_gpu.py -x -q > gpurun_out/f_tests.log 2>&1; echo "pytest rc $?" >> gpurun_out/f_tests.log
tail -4 gpurun_out/f_tests.log
run() { tag=$1; shift; env "$@" > gpurun_out/f_bench_$tag.json 2> gpurun_out/f_bench_$tag.err; python -c "
import json;d=json.loads(open('gpurun_out/f_bench_$tag.json').read().strip().splitlines()[-1]);print('$tag',d['ms_per_step'],d['value'],d['loss'],d['gpu_launches']//d['steps'])" || tail -5 gpurun_out/f_bench_$tag.err; }
run lora A=1 python bench.py --lora --no-cpu-baseline --no-gpu-eager-baseline --no-optimizer
run lora_noahead OF_LORA_MERGE_AHEAD=0 python bench.py --lora --no-cpu-baseline --no-gpu-eager-baseline --no-optimizer
run train_critpath OF_DEBUG_SKIP_OFFPATH=1 python bench.py --no-cpu-baseline --no-gpu-eager-baseline --no-optimizer
run lora_critpath OF_DEBUG_SKIP_OFFPATH=1 python bench.py --lora --no-cpu-baseline --no-gpu-eager-baseline --no-optimizer
