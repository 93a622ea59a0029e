#!/bin/bash
# DDP tuning at N GPUs: train bench only, variants of NCCL CTA budget / registered arena.  usage: gpu_call_ddp8.sh N
N=${1:-8}
mkdir -p gpurun_out
tr() { tag=$1; port=$2; shift 2; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 10 --warmup 3 --no-comm-breakdown --no-optimizer $EXTRA > gpurun_out/ddp${N}_$tag.json 2> gpurun_out/ddp${N}_$tag.err; python -c "
import json;d=json.loads(open('gpurun_out/ddp${N}_$tag.json').read().strip().splitlines()[-1]);print('$tag',round(d['ms_per_step'],2),round(d['value'],1))" || tail -n 5 gpurun_out/ddp${N}_$tag.err; }
EXTRA="--reserve-sms 16" tr reg16 29521 NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,REG,TUNING NCCL_DEBUG_FILE=gpurun_out/nccl_${N}_%p.log
cat gpurun_out/nccl_${N}_*.log | grep "AllReduce: [0-9]\{8,\}" | cut -c1-160 | sort | uniq -c | sort -rn | head -4; cat gpurun_out/nccl_${N}_*.log | grep -i "register comm" | head -2; rm -f gpurun_out/nccl_${N}_*.log
EXTRA="--reserve-sms 16" tr plain16 29522 OF_DDP_REGISTERED_ARENA=0
EXTRA="--reserve-sms 32" tr reg32 29524 A=1
EXTRA="--reserve-sms 32" tr plain32 29525 OF_DDP_REGISTERED_ARENA=0
