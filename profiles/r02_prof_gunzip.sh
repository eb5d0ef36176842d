timeout 300 python -m pytest tests/test_gpu_gunzip.py -x -q -m gpu -k cli 2>&1 | grep -E "^E" | head -40 > gpurun_out/gunzip_cli_fail.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gunzip_kernel -c 1 -o gpurun_out/r02_gunzip_v1 python profiles/r02_gunzip_perf.py 296 > gpurun_out/ncu_gunzip.log 2>&1
echo done
