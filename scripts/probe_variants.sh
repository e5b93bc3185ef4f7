#!/bin/bash
# scripts/probe_variants.sh <tag> <variant>...: quick_probe (1e9-key index, 1 M pairs) for the in-tree library and each variants/<v>
tag=$1; shift
timeout 200 python scripts/quick_probe.py 2.5e6 1e6 > gpurun_out/${tag}_main.log 2>&1
for v in "$@"; do
  UMGAP_GPU_LIB=$PWD/variants/$v/libumgap_gpu.so timeout 200 python scripts/quick_probe.py 2.5e6 1e6 > gpurun_out/${tag}_$v.log 2>&1
done
grep -H "classify (both" gpurun_out/${tag}_*.log
