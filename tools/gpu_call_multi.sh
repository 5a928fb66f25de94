#!/bin/bash
# One multi-GPU call (gpurun --gpus N): the C multi-device test, the host->device probe and the bench line at N ranks.
N=${1:-2}; tag=${2:-multi}
mkdir -p gpurun_out
tests/c/bin/test_multi > gpurun_out/${tag}_test_multi.log 2>&1; echo "test_multi rc=$?"; tail -4 gpurun_out/${tag}_test_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/h2d_probe.py > gpurun_out/${tag}_h2d.json 2> gpurun_out/${tag}_h2d.err; echo "h2d rc=$?"; cat gpurun_out/${tag}_h2d.json | cut -c1-1500
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${tag}_bench.err
python tools/bench_summary.py gpurun_out/${tag}_bench.json
