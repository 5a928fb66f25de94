#!/bin/bash
# ncu --set full capture of one kernel of a small full-pipeline batch; usage: tools/prof_env.sh <kernel-regex> <out-name>
ncu --set full --clock-control none --import-source on -k regex:"$1" -s 1 -c 1 -o gpurun_out/$2 python bench.py --steps 1 --warmup 1 --songs-per-step 128 --no-cpu --no-spectral --no-distance --e2e-songs 2 > gpurun_out/b_ncu2.log 2>&1
tail -1 gpurun_out/b_ncu2.log | cut -c1-100
