#!/bin/bash
# round 2: builds variants of the sketch hash kernel (k = 21, 32, 16 objects only; everything else is shared) as separate
# shared libraries under fp-mash_b200/build/variants/, and -- with "run" -- times each on the GPU (profiles/sketch_quick.py).
# usage: bash profiles/r02_sketch_variants.sh build|run  "name=flags" ...
set -e
MODE=$1; shift
cd "$(dirname "$0")/../fp-mash_b200"
NVF="-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC"
OTHERS=$(ls build/*.o | grep -v "sketch_k21.o\|sketch_k32.o\|sketch_k16.o")
if [ "$MODE" = build ]; then
  rm -rf build/variants; mkdir -p build/variants
  for spec in "$@"; do
    v=${spec%%=*}; flags=${spec#*=}
    echo "$v=$flags" >> build/variants/LIST
    ( for k in 21 32 16; do nvcc $NVF $flags -DFPM_K=$k -c csrc/sketch_inst.cu -o build/variants/${v}_k$k.o; done
      nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/lib_$v.so $OTHERS build/variants/${v}_k21.o build/variants/${v}_k32.o build/variants/${v}_k16.o -lcudart
      rm -f build/variants/${v}_k*.o ) &
  done
  wait
  ls build/variants/
else
  cd ..
  while IFS= read -r spec; do
    v=${spec%%=*}
    echo "== $spec"
    FPMASH_B200_LIB=$PWD/fp-mash_b200/build/variants/lib_$v.so python profiles/sketch_quick.py --all 2>&1 | tail -3
  done < fp-mash_b200/build/variants/LIST
fi
