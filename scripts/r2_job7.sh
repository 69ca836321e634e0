#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/j7_pytest.log 2>&1; tail -8 gpurun_out/j7_pytest.log
python scripts/time_schemes.py 512 > gpurun_out/j7_schemes_512.txt 2>&1; cat gpurun_out/j7_schemes_512.txt
