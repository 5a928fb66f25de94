#!/bin/bash
# ncu --set full captures of the two float32 pass1 variants (FULL and spectral-only); usage: tools/prof_p1.sh <tag>
B="python bench.py --steps 1 --warmup 1 --songs-per-step 128 --no-cpu --no-distance --e2e-songs 2"
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:pass1_kernelILi2ELb1 -s 1 -c 1 -o gpurun_out/p1_full_$1 -f $B > gpurun_out/p1_full.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:pass1_kernelILi2ELb0 -s 1 -c 1 -o gpurun_out/p1_lite_$1 -f $B > gpurun_out/p1_lite.log 2>&1
tail -2 gpurun_out/p1_lite.log | cut -c1-200
