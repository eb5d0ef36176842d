# launch lists of one rank's dist step at the block shapes of 1, 2 and 8 GPUs (configs[2])
for shape in "20000 20000" "20000 10000" "10000 5000"; do
  set -- $shape
  timeout 200 python profiles/r02_dist_block_prof.py $1 $2
  timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_dist_launches_$1x$2.csv python profiles/r02_dist_block_prof.py $1 $2 > /dev/null 2>&1
done
