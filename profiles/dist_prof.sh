# launch lists (per-kernel durations) of one dist step: matrix mode and hits mode
for m in matrix hits; do
  timeout 200 python profiles/dist_prof.py $m
  timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/dist_launches_$m.csv python profiles/dist_prof.py $m > /dev/null 2>&1
done
