#!/bin/bash
# Run on the GPU box: ncu launch list of a short full-pipeline bench + --set full captures of every kernel.
# Outputs land in gpurun_out/ (r1_launches.csv, r1_*.ncu-rep); tools/summarise_ncu.py turns them into profiles/*.txt.
set -x
B="python bench.py --steps 2 --warmup 1 --songs-per-step 128 --no-cpu --e2e-songs 4 --distance-vectors 65536"
BL="python bench.py --steps 2 --warmup 1 --songs-per-step 512 --no-cpu --e2e-songs 32 --distance-vectors 262144"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pass1|epilogue|envelope|tail|logcomp|distance" -c 80 --csv --log-file gpurun_out/r1_launches.csv $BL > gpurun_out/r1_launches.log 2>&1
for k in pass1_kernel envelope_kernel tail_kernel epilogue_kernel logcomp_kernel distance_nearest_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:"$k" -s 1 -c 1 -o gpurun_out/r1_$k -f $B > gpurun_out/r1_$k.log 2>&1
done
# the lite (spectral-only) pass1 variant: third pass1 launch family; capture the last pass1 launch of the run
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:pass1_kernelILi2ELb0 -s 1 -c 1 -o gpurun_out/r1_pass1_lite -f $B > gpurun_out/r1_pass1_lite.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit --format=csv > gpurun_out/r1_smi.txt
ls -la gpurun_out
