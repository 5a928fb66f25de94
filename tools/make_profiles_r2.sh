#!/bin/bash
# Run on the GPU box (gpurun): ncu launch list of the default bench command + --set full captures of every kernel of a
# full-pipeline step AT THE BENCHMARK BATCH SIZE (2 048 songs; spectral-only: 1 024 x 30 s). Outputs in gpurun_out/;
# tools/summarise_ncu.py turns the .ncu-rep files into profiles/r2_*.txt here afterwards.
set -x
mkdir -p gpurun_out
SHORT="--no-cpu --parity-songs 0 --chain-songs 0 --no-bl-analyze --e2e-songs 8 --s16-songs 256"
# launch list: 2 timed steps of the default workload (2 048 songs), all legs that launch kernels
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pass1|epilogue|envelope|tail|logcomp|distance|nearest|resample|dfma" \
    -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 $SHORT --distance-vectors 262144 > gpurun_out/r2_launches.log 2>&1
B="python bench.py --steps 1 --warmup 1 --songs-per-step 2048 $SHORT --no-distance"
for k in envelope_kernel pass1_kernel tail_kernel epilogue_kernel logcomp_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:"$k" -s 4 -c 1 -f -o gpurun_out/r2_$k $B --no-spectral > gpurun_out/r2_$k.log 2>&1
done
# the spectral-only (lite) pass1 at its benchmark size: the last pass1 launches of the run belong to that leg
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:pass1_kernelILi2ELb0 -s 2 -c 1 -f -o gpurun_out/r2_pass1_lite $B > gpurun_out/r2_pass1_lite.log 2>&1
# native int16 stereo instantiations (envelope_kernel<false>, pass1<S16, FULL>)
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:envelope_kernelILb0 -s 1 -c 1 -f -o gpurun_out/r2_envelope_s16 $B --no-spectral --s16-songs 1024 > gpurun_out/r2_envelope_s16.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:distance_nearest -s 1 -c 1 -f -o gpurun_out/r2_distance_nearest python bench.py --steps 1 --warmup 1 --songs-per-step 64 $SHORT --no-spectral --s16-songs 0 > gpurun_out/r2_distance_nearest.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit --format=csv > gpurun_out/r2_smi.txt
ls -la gpurun_out | tail -20
