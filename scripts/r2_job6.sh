#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/j6_pytest.log 2>&1; tail -15 gpurun_out/j6_pytest.log
python scripts/time_schemes.py 512 > gpurun_out/j6_schemes_512.txt 2>&1; cat gpurun_out/j6_schemes_512.txt
python scripts/time_schemes.py 256 > gpurun_out/j6_schemes_256.txt 2>&1
