#!/bin/bash
# ncu evidence for the bench workload: launch list + full capture of the hot kernels.
# Usage (on the GPU box): scripts/profile_bench.sh <tag>
tag=${1:-r02}
args="--steps 2 --warmup 1 --no-cpu-baseline --no-e2e --legs none"
python bench.py $args > gpurun_out/plain_${tag}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${tag}.csv python bench.py $args > gpurun_out/ncu_launches_${tag}.log 2>&1
python bench.py $args > gpurun_out/plain2_${tag}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'lookup_sampled|classify_kernel' -s 2 -c 2 -o gpurun_out/prof_${tag} python bench.py $args > gpurun_out/ncu_full_${tag}.log 2>&1
tail -2 gpurun_out/ncu_full_${tag}.log
