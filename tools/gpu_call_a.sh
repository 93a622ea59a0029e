#!/bin/bash
# Round-2 state capture on one B200: GPU tests, the driver's bench line, launch list of one graph replay, --set full of the
# attention / GEMM kernels, configs 4 and 5 at one GPU.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02_gputests.log
tail -3 gpurun_out/r02_gputests.log
python bench.py > gpurun_out/r02_bench_train_cfgL_1gpu.json 2> gpurun_out/r02_bench_train_cfgL_1gpu.err; tail -c 600 gpurun_out/r02_bench_train_cfgL_1gpu.json
bash tools/ncu_step_list.sh gpurun_out/r02_ncu_launches_one_graph_replay.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attn_|gemm_kernel' -c 12 -f -o gpurun_out/r02_attn_gemm_full python tools/ncu_one.py > gpurun_out/r02_ncu_full.log 2>&1
python bench.py --lora --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/r02_bench_lora_cfgL_1gpu.json 2> gpurun_out/r02_bench_lora.err; tail -c 300 gpurun_out/r02_bench_lora_cfgL_1gpu.json
python bench.py --mode sample --size S --frames 32768 --batch 1 --steps 2 > gpurun_out/r02_bench_sample_cfgS_32768_1gpu.json 2> gpurun_out/r02_bench_sample.err; tail -c 300 gpurun_out/r02_bench_sample_cfgS_32768_1gpu.json
