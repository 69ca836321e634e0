#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/j17_pytest.log 2>&1; tail -6 gpurun_out/j17_pytest.log
python scripts/time_gradient.py 512 256 1024 > gpurun_out/j17_gradient.txt 2>&1; cat gpurun_out/j17_gradient.txt
python bench.py --steps 50 --warmup 5 > gpurun_out/j17_bench.json 2> gpurun_out/j17_bench.err; echo "bench rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
