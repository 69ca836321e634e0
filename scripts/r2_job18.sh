#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "out_of_bounds or two_streams" > gpurun_out/j18_pytest.log 2>&1; tail -3 gpurun_out/j18_pytest.log
python scripts/time_gradient.py 512 256 1024 128 > gpurun_out/j18_gradient.txt 2>&1; cat gpurun_out/j18_gradient.txt
