# round 2, final captures: launch list of the default bench shape (reduced sizes) + ncu --set full of the dominant kernels
CMD="python bench.py --steps 2 --warmup 1 --genomes 60 --dist-sketches 3200 --no-cpu --no-configs"
$CMD > gpurun_out/r02_prof_plain.log 2>&1 || { tail -5 gpurun_out/r02_prof_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_l.log 2>&1
bash profiles/r02_prof_sketch.sh v3
for k in dist_tile32_kernel dist_qcode_kernel dist_mark_kernel dist_fill_unshared_kernel; do
  timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k -c 1 -f -o gpurun_out/r02_$k python profiles/dist_prof.py > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep | tail -8
