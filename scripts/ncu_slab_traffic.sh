# gpurun job behind profiles/r2x_ncu_dram_xy_*: DRAM bytes + duration of the fused x/y launch on the slab shapes (final draw order, and 16-tile squares)
for nz in 128 256 512; do
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:stream_kernel_xy -s 3 -c 5 --csv --log-file gpurun_out/y_dram_xy_${nz}.csv python scripts/time_xy_one.py $nz 1024 1024 2 > /dev/null 2>&1
done
CFD_XY_SUB=16 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:stream_kernel_xy -s 3 -c 5 --csv --log-file gpurun_out/y_dram_xy_128_squares.csv python scripts/time_xy_one.py 128 1024 1024 2 > /dev/null 2>&1
for f in gpurun_out/y_dram_xy_*.csv; do echo $f; grep -o '"dram__bytes_read.sum","byte","[0-9]*"\|"dram__bytes_write.sum","byte","[0-9]*"\|"gpu__time_duration.sum","ns","[0-9]*"' $f | tr '\n' ' '; echo; done
