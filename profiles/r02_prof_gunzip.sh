# ncu capture of the gunzip kernel (296 files x 5 MB): bash profiles/r02_prof_gunzip.sh [tag]
tag=${1:-v2}
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gunzip_kernel -c 1 -o gpurun_out/r02_gunzip_$tag python profiles/r02_gunzip_perf.py 296 > gpurun_out/ncu_gunzip.log 2>&1
echo done
