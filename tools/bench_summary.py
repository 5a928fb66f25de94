"""Prints the numbers of one bench.py line that matter when iterating on kernels."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("songs/s", round(d["value"]), "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"])
print("roofline", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d["roofline"].items() if k in ("kernel", "achieved", "peak", "frac")})
for k, v in d["roofline_kernels"].items():
    print("  %-16s %8.3f ms/step  share %.3f  hbm %.3f  fp64 %.3f" % (k, v.get("ms_per_step", v["ms_per_launch"]), v["share"], v["frac_hbm"], v.get("frac_fp64", 0)))
for key in ("spectral_only", "native_s16", "all_pairs", "cpu_baseline", "parity", "parity_sample", "configs4_chained", "bl_analyze_path", "clocks"):
    if d.get(key):
        v = dict(d[key])
        v.pop("workload", None); v.pop("sample", None); v.pop("mix", None)
        print(key, json.dumps(v)[:900])
