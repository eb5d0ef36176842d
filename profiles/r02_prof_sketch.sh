# round 2: ncu --set full of the sketch hash kernel (k = 21), run through profiles/sketch_quick.py (100 x 5 Mbp resident)
# usage: bash profiles/r02_prof_sketch.sh <tag> [lib]     -> gpurun_out/r02_sketch_hash_<tag>.ncu-rep
TAG=${1:-v1}
[ -n "$2" ] && export FPMASH_B200_LIB=$2
python profiles/sketch_quick.py > gpurun_out/r02_quick_$TAG.log 2>&1 || { tail -5 gpurun_out/r02_quick_$TAG.log; exit 1; }
cat gpurun_out/r02_quick_$TAG.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sketch_hash_kernel_v2 -s 3 -c 1 -f -o gpurun_out/r02_sketch_hash_$TAG python profiles/sketch_quick.py > gpurun_out/ncu_sk_$TAG.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
