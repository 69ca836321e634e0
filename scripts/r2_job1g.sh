#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/j26_pytest.log 2>&1; tail -4 gpurun_out/j26_pytest.log
for nz in 128 256 512; do
  ZSTEP_ONLY="zx (" ncu --set full --clock-control none -k regex:"stream_kernel_zx|stream_kernel_xy" -s 6 -c 2 -o gpurun_out/j26_ncu_zx_$nz python scripts/time_zpart_step.py $nz 1024 3 > gpurun_out/j26_ncu_zx_$nz.log 2>&1
  ncu -i gpurun_out/j26_ncu_zx_$nz.ncu-rep --page raw --csv > gpurun_out/j26_ncu_zx_${nz}_raw.csv 2>/dev/null; rm -f gpurun_out/j26_ncu_zx_$nz.ncu-rep
done
python scripts/time_zpart_step.py 512 1024 10 > gpurun_out/j26_zstep_512.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
