#!/bin/bash
# Final single-GPU verification of the round: all GPU tests, smoke(), the driver's default bench line, launch list of one graph replay
# (with DRAM bytes), then compute-sanitizer on the tiny configuration.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_final_gputests.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02_final_gputests.log
tail -3 gpurun_out/r02_final_gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1; echo "smoke rc $?" >> gpurun_out/r02_final_smoke.log; tail -3 gpurun_out/r02_final_smoke.log
python bench.py > gpurun_out/r02_final_bench_train_cfgL_1gpu.json 2> gpurun_out/r02_final_bench.err; python -c "
import json;d=json.loads(open('gpurun_out/r02_final_bench_train_cfgL_1gpu.json').read().strip().splitlines()[-1]);print('bench',d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['frac'],d['gpu_eager_baseline'],d['clocks'])"
python bench.py --lora --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/r02_final_bench_lora_cfgL_1gpu.json 2> gpurun_out/r02_final_lora.err; python -c "
import json;d=json.loads(open('gpurun_out/r02_final_bench_lora_cfgL_1gpu.json').read().strip().splitlines()[-1]);print('lora',d['ms_per_step'],d['value'])"
OF_PROFILE_STEP=1 timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_final_launches_hbm.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline --no-optimizer > gpurun_out/r02_final_ncu_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_final_launches_hbm.csv 12
SANITIZER_TIMEOUT=240 bash tools/sanitize.sh gpurun_out
