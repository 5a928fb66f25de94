#!/bin/bash
# quick GPU loop: parity tests + a short bench summary
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 4 --warmup 3 --no-cpu "$@" > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; tail -2 gpurun_out/bench_q.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_q.json"))
print("songs/s", round(d["value"]), "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"]))
for k, v in d["roofline_kernels"].items():
    print(" ", k, round(v["ms_per_launch"], 3), "ms", round(v["share"], 3), "hbm", round(v["frac_hbm"], 3), "fp64", round(v.get("frac_fp64", 0), 3))
if d.get("all_pairs"): print("all_pairs", d["all_pairs"]["ms"], "ms", "%.3g pairs/s" % d["all_pairs"]["pairs_per_s"], d["all_pairs"]["spot_check_ok"], "gather ms", d["all_pairs"]["all_gather_ms"])
if d.get("native_s16"): print("native s16", round(d["native_s16"]["value"]), d["native_s16"]["ms_per_pass"], d["native_s16"]["all_status_ok"])
if d.get("spectral_only"): print("spectral frac", round(d["spectral_only"]["roofline"]["frac"], 3), d["spectral_only"]["value"])
PY
