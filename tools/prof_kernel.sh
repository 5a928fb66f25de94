#!/bin/bash
# ncu --set full capture of ONE kernel of a full-pipeline step at a given batch size (default: the benchmark's 2048 songs)
# usage: tools/prof_kernel.sh <kernel-regex> <out-name> [songs-per-step] [extra bench args]
k=$1; out=$2; songs=${3:-2048}; shift; shift; shift
ncu --set full --clock-control none --import-source on -k regex:"$k" -s 1 -c 1 -f -o gpurun_out/$out \
  python bench.py --steps 1 --warmup 1 --songs-per-step $songs --no-cpu --no-spectral --no-distance --e2e-songs 2 --s16-songs 0 \
  --parity-songs 0 --chain-songs 0 --no-bl-analyze "$@" > gpurun_out/${out}_ncu.log 2>&1
tail -2 gpurun_out/${out}_ncu.log | cut -c1-200
