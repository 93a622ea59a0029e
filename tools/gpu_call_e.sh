#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lora_gpu.py tests/test_kernels_gpu.py tests/test_model_parity_gpu.py tests/test_optim_gpu.py -x -q > gpurun_out/e_tests.log 2>&1; echo "pytest rc $?" >> gpurun_out/e_tests.log
tail -4 gpurun_out/e_tests.log
run() { tag=$1; shift; env "$@" > gpurun_out/e_bench_$tag.json 2> gpurun_out/e_bench_$tag.err; python -c "
import json;d=json.loads(open('gpurun_out/e_bench_$tag.json').read().strip().splitlines()[-1]);print('$tag',d['ms_per_step'],d['value'],d['loss'],d['gpu_launches']//d['steps'])" || tail -5 gpurun_out/e_bench_$tag.err; }
run train A=1 python bench.py --no-cpu-baseline --no-gpu-eager-baseline --no-optimizer
run train_bnlegacy OF_GEMM_BN_LEGACY=1 python bench.py --no-cpu-baseline --no-gpu-eager-baseline --no-optimizer
run lora A=1 python bench.py --lora --no-cpu-baseline --no-gpu-eager-baseline
run lora_ungrouped OF_LORA_GROUPED=0 python bench.py --lora --no-cpu-baseline --no-gpu-eager-baseline --no-optimizer
