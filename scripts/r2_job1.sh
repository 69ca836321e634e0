#!/bin/bash
# round-2 GPU job 1: tests, bench, knob sweeps (one B200)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/j1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j1_pytest.log
tail -5 gpurun_out/j1_pytest.log
python bench.py --steps 50 --warmup 5 > gpurun_out/j1_bench.json 2> gpurun_out/j1_bench.err; echo "bench rc=$?"
python scripts/sweep_256.py 256 > gpurun_out/j1_sweep256.txt 2>&1
XY_ACTIVES=default XY_WARPS=6,5,7 python scripts/time_xy.py 128 1024 1024 > gpurun_out/j1_xy_slab.txt 2>&1
for sub in 8 12 16 32; do CFD_XY_SUB=$sub XY_ACTIVES=default,6,10,16 XY_WARPS=6 python scripts/time_xy.py 128 1024 1024 >> gpurun_out/j1_xy_slab.txt 2>&1; done
XY_ACTIVES=default XY_WARPS=6 XY_TAUS=,500,1000,2000 python scripts/time_xy.py 256 1024 1024 >> gpurun_out/j1_xy_slab.txt 2>&1
tail -3 gpurun_out/j1_sweep256.txt
