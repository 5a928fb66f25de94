#!/bin/bash
# Builds variants of libbliss.so that differ in the envelope kernel's build-time switches (on the CPU box, before
# gpurun) or times them (on the GPU box): tools/ab_envelope.sh build | run [songs]
# Variants live in tools/variants/ (git-ignored, shipped by gpurun).
set -e
cd "$(dirname "$0")/.."
VAR=tools/variants
declare -A FLAGS=(
  [base]=""
  [pshfl]="-DBLX_ENV_PARTNER_SHFL=1"
)
if [ "$1" = build ]; then
  mkdir -p $VAR
  make -C bliss_b200 -j8 >/dev/null
  for v in "${!FLAGS[@]}"; do
    for src in envelope pass1; do
      nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC ${FLAGS[$v]} $EXTRA -Xptxas -v \
        -c bliss_b200/csrc/$src.cu -o $VAR/${src}_$v.o 2> $VAR/ptxas_${src}_$v.txt
    done
    cp $VAR/ptxas_envelope_$v.txt $VAR/ptxas_$v.txt
    objs=$(ls bliss_b200/csrc/*.o bliss_b200/host/*.o | grep -v -e csrc/envelope.o -e csrc/pass1.o)
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $VAR/libbliss_$v.so $objs $VAR/envelope_$v.o $VAR/pass1_$v.o -ldl -lpthread -lm
    echo "$v: $(grep -A1 'envelope_kernelILb1' $VAR/ptxas_$v.txt | grep -o 'bytes spill stores' | head -1) $(grep -A2 'envelope_kernelILb1' $VAR/ptxas_$v.txt | grep -o 'Used [0-9]* registers' | head -1)"
  done
else
  songs=${2:-1024}
  for so in $VAR/libbliss_*.so; do
    v=$(basename $so .so)
    BLX_DEBUG_OCCUPANCY=1 BLISS_B200_LIB=$PWD/$so python bench.py --steps 3 --warmup 2 --songs-per-step $songs --no-cpu --no-distance \
      --e2e-songs 2 --s16-songs 128 --parity-songs 0 --chain-songs 0 --no-bl-analyze > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err || echo "$v FAILED"
    grep -h "CTAs/SM" gpurun_out/ab_$v.err | head -1
    python - "$v" gpurun_out/ab_$v.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
k = d["roofline_kernels"]
sp = d.get("spectral_only") or {}
print("%-18s env %.3f ms  pass1 %.3f  step %.2f ms  native_s16 %.0f songs/s  spectral %.0f songs/s frac %.3f" % (sys.argv[1],
      k["envelope_kernel"]["ms_per_step"], k["pass1_kernel"]["ms_per_step"], d["ms_per_step"],
      d["native_s16"]["value"] if d.get("native_s16") else 0, sp.get("value", 0), (sp.get("roofline") or {}).get("frac", 0)))
PY
  done
fi
