#!/bin/bash
# One GPU call: the -m gpu suite, then the default bench line (all legs), each logged under gpurun_out/.
# usage: tools/gpu_call.sh <tag> [bench args...]
tag=${1:-call}; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
nproc >> gpurun_out/${tag}_smi.txt; free -g >> gpurun_out/${tag}_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${tag}_pytest.log
timeout 1500 python bench.py "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${tag}_bench.err
python tools/bench_summary.py gpurun_out/${tag}_bench.json
