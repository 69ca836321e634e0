# gpurun job behind profiles/r2x_bench_launches.csv, r2x_ncu_full_raw_*: each ncu pass runs only after the same command exited 0 without ncu
set -x
python bench.py --steps 5 --warmup 3 --no-cpu --no-extra > gpurun_out/x_bench_plain.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/x_bench_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-extra > gpurun_out/x_ncu1.log 2>&1
python scripts/prof_gradient.py 512 3 > gpurun_out/x_prof_plain.txt 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 2 -c 2 -f -o gpurun_out/x_full_512 python scripts/prof_gradient.py 512 3 > gpurun_out/x_ncu2.log 2>&1
ncu -i gpurun_out/x_full_512.ncu-rep --page raw --csv > gpurun_out/x_full_512_raw.csv 2>/dev/null
ZSTEP_ONLY="zx (zpart" python scripts/time_zpart_step.py 128 1024 3 > gpurun_out/x_zstep_plain.txt 2>&1 && \
ZSTEP_ONLY="zx (zpart" ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 4 -c 2 -f -o gpurun_out/x_full_slab python scripts/time_zpart_step.py 128 1024 3 > gpurun_out/x_ncu3.log 2>&1
ncu -i gpurun_out/x_full_slab.ncu-rep --page raw --csv > gpurun_out/x_full_slab_raw.csv 2>/dev/null
ls -la gpurun_out/ | tail -15
rm -f gpurun_out/*.ncu-rep
