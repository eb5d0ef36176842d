CMD="python bench.py --steps 1 --warmup 1 --genomes 20 --dist-sketches 3200 --no-cpu"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:dist_tile32_kernel -s 1 -c 1 -o gpurun_out/r01_dist_tile32_v2 -f $CMD > gpurun_out/ncu4.log 2>&1
ls -la gpurun_out | tail -2
