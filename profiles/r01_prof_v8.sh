# round-1 eighth capture: the pruned dist step (configs[2] shape): launch lists of both modes + ncu --set full of the tile kernel
# (pruned + grouped), the marking kernel and the union-find pass.  Summaries: python profiles/ncu_summary.py <rep> > profiles/<name>.txt
bash profiles/dist_prof.sh || exit 1
for k in dist_tile32_kernel dist_mark_kernel dist_uf_union_kernel; do
  timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k -c 1 -f -o gpurun_out/r01_${k}_v8 python profiles/dist_prof.py > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out | tail -5
