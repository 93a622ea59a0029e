#!/bin/bash
# OUTCOME: every variant failed with "no algorithm/protocol available for AllReduce ... NCCL_ALGO was set to allreduce:nvls" because the
# all-reduce used ReduceOp.AVG (no NVLS form).  That finding led to mean-by-pre-division + SUM (ddp.py); see gpu_call_ddp8c.sh.
mkdir -p gpurun_out
tr() { N=$1; tag=$2; port=$3; shift 3; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 10 --warmup 3 --no-comm-breakdown --no-optimizer $EXTRA > gpurun_out/ddpb${N}_$tag.json 2> gpurun_out/ddpb${N}_$tag.err; python -c "
import json;d=json.loads(open('gpurun_out/ddpb${N}_$tag.json').read().strip().splitlines()[-1]);print('$N $tag',round(d['ms_per_step'],2),round(d['value'],1))" || tail -n 8 gpurun_out/ddpb${N}_$tag.err | cut -c1-300; }
EXTRA="--reserve-sms 16" tr 8 nvls16 29531 NCCL_ALGO=allreduce:nvls NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=TUNING NCCL_DEBUG_FILE=gpurun_out/nccl_%p.log
cat gpurun_out/nccl_*.log | grep "AllReduce: [0-9]\{8,\}" | sed 's/.*AllReduce: [0-9]* Bytes/AllReduce/' | sort | uniq -c | sort -rn | head -4; rm -f gpurun_out/nccl_*.log
EXTRA="--reserve-sms 8" tr 8 nvls8 29532 NCCL_ALGO=allreduce:nvls
EXTRA="--reserve-sms 32" tr 8 nvls32 29533 NCCL_ALGO=allreduce:nvls
EXTRA="--reserve-sms 16" tr 4 nvls16 29534 NCCL_ALGO=allreduce:nvls
