#!/bin/bash
# 8 GPUs: record run with the one-kernel d/dz
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581"
timeout 400 $TR bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/j25_bench_n8.json 2> gpurun_out/j25_bench_n8.err; echo "bench rc=$?"
timeout 300 $TR scripts/time_zpart_step_mp.py 128 1024 50 > gpurun_out/j25_step_mp_128.txt 2>&1; echo "step rc=$?"
grep -v "Warning\|^\*\*\*\|OMP" gpurun_out/j25_step_mp_128.txt | tail -14
timeout 600 $TR scripts/check_partition_nccl.py 1024 > gpurun_out/j25_check_partition_p8.txt 2>&1; echo "check rc=$?"
grep -c " OK" gpurun_out/j25_check_partition_p8.txt; grep "FAIL\|Error" gpurun_out/j25_check_partition_p8.txt | head
python - <<'PY'
import json
d = json.load(open("gpurun_out/j25_bench_n8.json")); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["frac_of_copy_ceiling"], d["check"], d["roofline"]["launches"], d["gpu_launches"])
PY
