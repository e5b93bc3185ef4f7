#!/bin/bash
# Random-sector sweep: table size, load-instruction variant, sliding window.  Output: gpurun_out/randsector_sweep.log
out=gpurun_out/randsector_sweep.log
: > $out
for mib in 64 128 256 512 1024 2048 4096 8192 16384 32768 65536 131072; do ./bench/randsector $mib 200 0 >> $out 2>&1; done
./bench/randsector 1024 200 all >> $out 2>&1
./bench/randsector 16384 200 all >> $out 2>&1
for w in 128 256 512 1024 2048 4096 16384; do ./bench/randsector 131072 200 0 $w >> $out 2>&1; done
for w in 128 256 512 1024 2048; do ./bench/randsector 16384 200 0 $w >> $out 2>&1; done
