#!/bin/bash
mkdir -p gpurun_out
ZSTEP_ONLY="zx" python scripts/time_zpart_step.py 128 1024 30 > gpurun_out/j28_zx2.txt 2>&1
ZSTEP_ONLY="zx" python scripts/time_zpart_step.py 256 1024 20 >> gpurun_out/j28_zx2.txt 2>&1
cat gpurun_out/j28_zx2.txt
