#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 --no-extra > gpurun_out/j31_bench_n1.json 2> gpurun_out/j31_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/j31_bench_n1.json")); print(d["value"], d["e2e"])
PY
