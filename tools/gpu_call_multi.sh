#!/bin/bash
# usage: gpu_call_multi.sh N tag   -- configs 3 (train DP), 5 (LoRA DP) and 4 (sampling replicas) on N GPUs of one box
N=${1:-2}; TAG=${2:-r02}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_ddp_nccl_gpu.py -x -q > gpurun_out/${TAG}_ddp_nccl_test.log 2>&1; tail -2 gpurun_out/${TAG}_ddp_nccl_test.log; fi
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
tr 29511 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_train_cfgL_${N}gpu.json 2> gpurun_out/${TAG}_train_${N}.err; tail -c 1500 gpurun_out/${TAG}_bench_train_cfgL_${N}gpu.json
tr 29512 --lora --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_lora_cfgL_${N}gpu.json 2> gpurun_out/${TAG}_lora_${N}.err; tail -c 1300 gpurun_out/${TAG}_bench_lora_cfgL_${N}gpu.json
tr 29513 --mode sample --size S --frames 32768 --batch 1 --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_sample_cfgS_32768_${N}gpu.json 2> gpurun_out/${TAG}_sample_${N}.err; tail -c 600 gpurun_out/${TAG}_bench_sample_cfgS_32768_${N}gpu.json
