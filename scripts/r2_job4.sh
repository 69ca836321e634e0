#!/bin/bash
# 4 GPUs: real NVLink exchange
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531"
timeout 300 $TR scripts/time_zpart_step_mp.py 128 1024 50 > gpurun_out/j4_step_mp_128.txt 2>&1; echo "rc=$?"
cat gpurun_out/j4_step_mp_128.txt | grep -v Warning | tail -14
timeout 300 $TR scripts/time_zpart_step_mp.py 256 1024 30 > gpurun_out/j4_step_mp_256.txt 2>&1; echo "rc=$?"
timeout 400 $TR scripts/check_partition_nccl.py 1024 > gpurun_out/j4_check_partition_p4.txt 2>&1; echo "check rc=$?"
grep -c OK gpurun_out/j4_check_partition_p4.txt; grep FAIL gpurun_out/j4_check_partition_p4.txt
timeout 300 $TR bench.py --gpus 4 --steps 50 --warmup 5 > gpurun_out/j4_bench_n4.json 2> gpurun_out/j4_bench_n4.err; echo "bench rc=$?"
