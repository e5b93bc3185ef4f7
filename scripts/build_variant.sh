#!/bin/bash
# scripts/build_variant.sh <name> <extra nvcc flags...>: builds variants/<name>/libumgap_gpu.so for A/B timing
name=$1; shift
mkdir -p variants/$name/obj
for f in umgap_b200/csrc/*.cu; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c $f -o variants/$name/obj/$(basename $f .cu).o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/$name/libumgap_gpu.so variants/$name/obj/*.o -cudart shared
rm -rf variants/$name/obj
