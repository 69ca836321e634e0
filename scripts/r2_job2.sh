#!/bin/bash
mkdir -p gpurun_out
python scripts/time_zpart_step.py 128 1024 20 > gpurun_out/j2_zstep_128.txt 2>&1
python scripts/time_zpart_step.py 256 1024 20 > gpurun_out/j2_zstep_256.txt 2>&1
python scripts/time_zpart_step.py 512 1024 10 > gpurun_out/j2_zstep_512.txt 2>&1
cat gpurun_out/j2_zstep_128.txt
ZSTEP_ONLY=fused ncu --set full --clock-control none --import-source on -k regex:stream_kernel_xy -c 2 -s 6 -o gpurun_out/j2_ncu_xyedge_128 python scripts/time_zpart_step.py 128 1024 3 > gpurun_out/j2_ncu.log 2>&1
ncu -i gpurun_out/j2_ncu_xyedge_128.ncu-rep --page raw --csv > gpurun_out/j2_ncu_xyedge_128_raw.csv 2>/dev/null
tail -3 gpurun_out/j2_ncu.log
