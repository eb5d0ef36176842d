# round-1 fourth capture: launch list of the default bench command shape + ncu --set full of the two dominant kernels
CMD="python bench.py --steps 2 --warmup 1 --genomes 60 --dist-sketches 3200 --no-cpu"
$CMD > gpurun_out/prof_plain_v4.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_launches_v4.csv $CMD > gpurun_out/ncu_l3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sketch_hash_kernel_v2 -s 1 -c 1 -f -o gpurun_out/r01_sketch_hash_v4 $CMD > gpurun_out/ncu_s3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dist_tile32_kernel -s 1 -c 1 -f -o gpurun_out/r01_dist_tile32_v4 $CMD > gpurun_out/ncu_d3.log 2>&1
ls -la gpurun_out | tail -4
