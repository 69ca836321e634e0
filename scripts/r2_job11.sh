#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/j11_pytest.log 2>&1; tail -6 gpurun_out/j11_pytest.log
python scripts/sweep_solver.py > gpurun_out/j11_solver_sweep.txt 2>&1; cat gpurun_out/j11_solver_sweep.txt
cp gpurun_out/solver_sweep.jsonl gpurun_out/j11_solver_sweep.jsonl
