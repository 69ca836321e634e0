#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29591"
timeout 300 $TR scripts/time_zpart_step_mp.py 128 1024 50 > gpurun_out/j27_step_mp_128.txt 2>&1; echo "rc=$?"
grep -v "Warning\|^\*\*\*\|OMP" gpurun_out/j27_step_mp_128.txt | tail -16
timeout 300 $TR scripts/time_zpart_step_mp.py 256 1024 30 > gpurun_out/j27_step_mp_256.txt 2>&1; echo "rc=$?"
grep -v "Warning\|^\*\*\*\|OMP" gpurun_out/j27_step_mp_256.txt | grep "zx"
