#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "scheme or toeplitz or lookahead or solver" > gpurun_out/j14_pytest.log 2>&1; tail -5 gpurun_out/j14_pytest.log
SCHEME_SWEEP=1 python scripts/time_schemes.py 512 > gpurun_out/j14_schemes_512.txt 2>&1; cat gpurun_out/j14_schemes_512.txt
