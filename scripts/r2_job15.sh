#!/bin/bash
mkdir -p gpurun_out
python scripts/time_gradient.py 512 256 128 1024 > gpurun_out/j15_gradient.txt 2>&1; cat gpurun_out/j15_gradient.txt
python -m pytest tests -m gpu -q -x > gpurun_out/j15_pytest.log 2>&1; tail -3 gpurun_out/j15_pytest.log
