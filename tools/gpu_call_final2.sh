#!/bin/bash
# Final 2-GPU verification: NCCL gradient-equivalence test + the driver's default bench line under torchrun (with comm breakdown)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ddp_nccl_gpu.py -x -q -s > gpurun_out/r02_final_ddp_nccl_test.log 2>&1; grep -E "DDP_NCCL_OK|passed|failed" gpurun_out/r02_final_ddp_nccl_test.log | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_final_bench_train_cfgL_2gpu.json 2> gpurun_out/r02_final_train2.err
python -c "
import json;d=json.loads(open('gpurun_out/r02_final_bench_train_cfgL_2gpu.json').read().strip().splitlines()[-1]);print('train2',d['ms_per_step'],d['value'],{k:v for k,v in d['comm'].items() if k!='bucket_mb'})" || tail -n 20 gpurun_out/r02_final_train2.err | cut -c1-300
