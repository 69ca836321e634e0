#!/bin/bash
# 2 GPUs
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 500 $TR scripts/check_partition_nccl.py 1024 > gpurun_out/j9_check_partition_p2.txt 2>&1; echo "check rc=$?"
grep -c " OK" gpurun_out/j9_check_partition_p2.txt; grep "FAIL\|Error" gpurun_out/j9_check_partition_p2.txt | head
timeout 400 $TR bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/j9_bench_n2.json 2> gpurun_out/j9_bench_n2.err; echo "bench rc=$?"
timeout 400 $TR bench.py --gpus 2 --steps 20 --warmup 5 --e2e pipelined --no-cpu > gpurun_out/j9_bench_n2_pipelined.json 2> gpurun_out/j9_bench_n2_pipelined.err; echo "bench pipelined rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/j9_bench_n2.json", "gpurun_out/j9_bench_n2_pipelined.json"):
    try:
        d = json.load(open(f)); print(f, d["value"], d["ms_per_step"], d["e2e"])
    except Exception as e:
        print(f, "ERR", e)
PY
