#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "zpart" > gpurun_out/j23_pytest_zpart.log 2>&1; tail -3 gpurun_out/j23_pytest_zpart.log
for w in 7 6; do
  echo "== zx warps=$w" >> gpurun_out/j23_zx.txt
  CFD_ZX_WARPS=$w ZSTEP_ONLY="zx" python scripts/time_zpart_step.py 128 1024 20 2>&1 | grep "zx" >> gpurun_out/j23_zx.txt
done
python scripts/time_zpart_step.py 128 1024 20 >> gpurun_out/j23_zx.txt 2>&1
python scripts/time_zpart_step.py 256 1024 20 >> gpurun_out/j23_zx.txt 2>&1
cat gpurun_out/j23_zx.txt
ZSTEP_ONLY="zx d/dz alone" ncu --set full --clock-control none -k regex:stream_kernel_zx -s 3 -c 1 -o gpurun_out/j23_ncu_zx python scripts/time_zpart_step.py 128 1024 3 > gpurun_out/j23_ncu.log 2>&1
ncu -i gpurun_out/j23_ncu_zx.ncu-rep --page raw --csv > gpurun_out/j23_ncu_zx_raw.csv 2>/dev/null; rm -f gpurun_out/j23_ncu_zx.ncu-rep
