#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/j3_pytest.log 2>&1; tail -3 gpurun_out/j3_pytest.log
python scripts/time_zpart_step.py 128 1024 20 > gpurun_out/j3_zstep_128.txt 2>&1
cat gpurun_out/j3_zstep_128.txt
ZSTEP_ONLY=fused ncu --set full --clock-control none --import-source on -k regex:stream_kernel_xy -s 2 -c 1 -o gpurun_out/j3_ncu_xyedge_128 python scripts/time_zpart_step.py 128 1024 3 > gpurun_out/j3_ncu.log 2>&1
ncu -i gpurun_out/j3_ncu_xyedge_128.ncu-rep --page raw --csv > gpurun_out/j3_ncu_xyedge_128_raw.csv 2>/dev/null
ZSTEP_ONLY="xy alone" ncu --set full --clock-control none -k regex:stream_kernel_xy -s 2 -c 1 -o gpurun_out/j3_ncu_xy_128 python scripts/time_zpart_step.py 128 1024 3 >> gpurun_out/j3_ncu.log 2>&1
ncu -i gpurun_out/j3_ncu_xy_128.ncu-rep --page raw --csv > gpurun_out/j3_ncu_xy_128_raw.csv 2>/dev/null
tail -3 gpurun_out/j3_ncu.log
