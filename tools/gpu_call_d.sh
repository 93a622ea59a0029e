#!/bin/bash
mkdir -p gpurun_out
python tools/dump_gemm_tags.py > gpurun_out/gemm_tags.json 2> gpurun_out/gemm_tags.err; tail -2 gpurun_out/gemm_tags.err
run() { tag=$1; shift; env "$@" python bench.py --no-cpu-baseline --no-gpu-eager-baseline --no-optimizer > gpurun_out/d_bench_$tag.json 2> gpurun_out/d_bench_$tag.err; python -c "
import json;d=json.loads(open('gpurun_out/d_bench_$tag.json').read().strip().splitlines()[-1]);print('$tag',d['ms_per_step'],d['value'],d['loss'])"; }
run default A=1
run nofwd OF_FWD_SIDE=0
run default2 A=1
run nofwd2 OF_FWD_SIDE=0
