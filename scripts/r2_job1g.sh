#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/j29_pytest.log 2>&1; tail -3 gpurun_out/j29_pytest.log
python bench.py --steps 100 --warmup 5 > gpurun_out/j29_bench_n1.json 2> gpurun_out/j29_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/j29_bench_ref.json 2> gpurun_out/j29_bench_ref.err; echo "ref rc=$?"
