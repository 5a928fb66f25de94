#!/bin/bash
# compute-sanitizer over a small slice of the GPU suite (run on the GPU box): memcheck on the analysis + distance paths,
# racecheck and synccheck on the kernels that exchange through shared memory. Logs under gpurun_out/.
mkdir -p gpurun_out
SEL='golden_fixture or hop_counts or ragged_batch or extreme or silent_passages or distance_nearest_matches or resampler_kernel'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$SEL" > gpurun_out/san_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -4 gpurun_out/san_memcheck.log
SEL2='golden_fixture or hop_counts or silence_f32 or spectrum_per_bin'
[ -n "$SKIP_RACE" ] || timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$SEL2" > gpurun_out/san_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -4 gpurun_out/san_racecheck.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$SEL" > gpurun_out/san_synccheck.log 2>&1; echo "synccheck rc=$?"
tail -4 gpurun_out/san_synccheck.log
# the wide pass: memcheck, synccheck and initcheck over nearly the whole GPU suite (seconds each)
K="not twenty_minute and not large_ragged and not many_threads"
for t in memcheck synccheck initcheck; do
  timeout 1000 compute-sanitizer --tool $t --print-limit 20 --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_compat.py tests/test_decode.py -q -m gpu -k "$K" > gpurun_out/san_${t}_all.log 2>&1; echo "$t (wide) rc=$?"
  tail -2 gpurun_out/san_${t}_all.log
done
