#!/bin/bash
# scripts/probe_variants.sh <tag> <variant>...: quick_probe (1e9-key index, 1 M pairs) for the in-tree library and each variants/<v>
tag=$1; shift
PROBE_SLICES=${PROBE_SLICES:-1,8} timeout 200 python scripts/quick_probe.py 2.5e6 1e6 > gpurun_out/${tag}_main.log 2>&1
for v in "$@"; do
  UMGAP_GPU_LIB=$PWD/variants/$v/libumgap_gpu.so PROBE_SLICES=${PROBE_SLICES:-1,8} timeout 200 python scripts/quick_probe.py 2.5e6 1e6 > gpurun_out/${tag}_$v.log 2>&1
done
grep -H -E "slices|kernels" gpurun_out/${tag}_*.log
