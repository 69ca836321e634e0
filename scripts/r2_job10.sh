#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551"
timeout 500 $TR scripts/check_partition_nccl.py 1024 > gpurun_out/j10_check_partition_p2.txt 2>&1; echo "check rc=$?"
grep -c " OK" gpurun_out/j10_check_partition_p2.txt; grep "FAIL\|Error" gpurun_out/j10_check_partition_p2.txt | head
timeout 300 $TR scripts/time_zpart_step_mp.py 512 1024 20 > gpurun_out/j10_step_mp_512.txt 2>&1
grep -v "Warning\|^\*\*\*\|OMP" gpurun_out/j10_step_mp_512.txt | tail -12
