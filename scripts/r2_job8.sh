#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/j8_pytest.log 2>&1; tail -6 gpurun_out/j8_pytest.log
for i in 1 2 3; do python -m pytest tests/test_zpart_ipc.py -m gpu -q 2>&1 | tail -1; done
