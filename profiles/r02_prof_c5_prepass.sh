# ncu capture of the configs[4] query-code lookup and pack kernels (after the tile-wise lookups / row-major codes)
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:"dist_qcode_kernel|dist_pack_queries_kernel|dist_mark_kernel" -c 3 -o gpurun_out/r02_c5_prepass python profiles/r02_c5_prof.py > gpurun_out/ncu_c5_prepass.log 2>&1
echo done
