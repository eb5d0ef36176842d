CMD="python bench.py --steps 1 --warmup 1 --genomes 60 --dist-sketches 0 --no-cpu"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sketch_hash_kernel -s 1 -c 1 -o gpurun_out/r01_sketch_hash_v2 $CMD > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out | tail -3
