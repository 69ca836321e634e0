#!/bin/bash
mkdir -p gpurun_out
CFD_PDL=1 python -m pytest tests -m gpu -x -q > gpurun_out/j5_pytest_pdl.log 2>&1; tail -2 gpurun_out/j5_pytest_pdl.log
for pdl in 0 1; do
  echo "== CFD_PDL=$pdl" >> gpurun_out/j5_pdl.txt
  CFD_PDL=$pdl python scripts/time_shape.py 256 256 256 512 512 512 64 64 64 128 128 128 >> gpurun_out/j5_pdl.txt 2>&1
  CFD_PDL=$pdl python scripts/time_zpart_step.py 128 1024 20 >> gpurun_out/j5_pdl.txt 2>&1
  CFD_PDL=$pdl python scripts/prof_gradient.py 512 20 >> gpurun_out/j5_pdl.txt 2>&1
done
cat gpurun_out/j5_pdl.txt
