#!/bin/bash
mkdir -p gpurun_out
python scripts/sanitize_case.py > gpurun_out/j16_plain.log 2>&1; tail -2 gpurun_out/j16_plain.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_case.py > gpurun_out/j16_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/j16_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize_case.py > gpurun_out/j16_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/j16_racecheck.log
