set -x
CMD="python bench.py --steps 1 --warmup 1 --genomes 60 --dist-sketches 3200 --no-cpu"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sketch_hash_kernel -s 1 -c 1 -o gpurun_out/r01_sketch_hash $CMD > gpurun_out/ncu2.log 2>&1
$CMD > gpurun_out/prof_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:dist_tile_kernel -s 1 -c 1 -o gpurun_out/r01_dist_tile $CMD > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out; tail -3 gpurun_out/ncu2.log gpurun_out/ncu3.log
