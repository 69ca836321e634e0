#!/bin/bash
mkdir -p gpurun_out
# (1) launch list of the bench command itself (after it ran plain)
python bench.py --steps 5 --warmup 3 --no-cpu --no-extra > gpurun_out/j13_bench_plain.json 2> gpurun_out/j13_bench_plain.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/j13_bench_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-extra > gpurun_out/j13_ncu_bench.log 2>&1; echo "ncu bench rc=$?"
# (2) --set full of the gradient kernels at 512^3
python scripts/prof_gradient.py 512 3 > gpurun_out/j13_prof_gradient.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 6 -c 2 -o gpurun_out/j13_ncu_gradient_512 python scripts/prof_gradient.py 512 3 > gpurun_out/j13_ncu_gradient.log 2>&1
ncu -i gpurun_out/j13_ncu_gradient_512.ncu-rep --page raw --csv > gpurun_out/j13_ncu_gradient_512_raw.csv 2>/dev/null; rm -f gpurun_out/j13_ncu_gradient_512.ncu-rep
# (3) per-rank kernels of the partitioned step on the 8- / 4- / 2-GPU slab shapes (middle rank, one GPU)
for nz in 128 256 512; do
  ZSTEP_ONLY=fused ncu --set full --clock-control none -k regex:"stream_kernel|reduced_planes" -s 9 -c 3 -o gpurun_out/j13_ncu_zstep_$nz python scripts/time_zpart_step.py $nz 1024 3 > gpurun_out/j13_ncu_zstep_$nz.log 2>&1
  ncu -i gpurun_out/j13_ncu_zstep_$nz.ncu-rep --page raw --csv > gpurun_out/j13_ncu_zstep_${nz}_raw.csv 2>/dev/null; rm -f gpurun_out/j13_ncu_zstep_$nz.ncu-rep
done
# (4) the general kernel (alpha = 1/3 solve, compact6 derivative) at 512^3
cat > /tmp/prof_g.py <<'PY'
import sys, torch
sys.path.insert(0, "/root/repo")
import compact_finite_differences_b200 as C
f = torch.rand((512, 512, 512), dtype=torch.float64, device="cuda"); df = torch.empty_like(f)
for a in (0, 1):
    op = C.CompactFiniteDifferenceSolver(f.shape, 0.1, a, scheme="compact6")
    s = C.NearToeplitzSolver(f.shape, (1., 2., 1. / 3, 1., 1. / 3, 2., 1.), axis=a)
    for _ in range(2):
        op(f, df); s.solve(df)
torch.cuda.synchronize()
PY
python /tmp/prof_g.py && ncu --set full --clock-control none --import-source on -k regex:stream_kernel_g -s 4 -c 4 -o gpurun_out/j13_ncu_general_512 python /tmp/prof_g.py > gpurun_out/j13_ncu_general.log 2>&1
ncu -i gpurun_out/j13_ncu_general_512.ncu-rep --page raw --csv > gpurun_out/j13_ncu_general_512_raw.csv 2>/dev/null; rm -f gpurun_out/j13_ncu_general_512.ncu-rep
ls -la gpurun_out/j13_*
