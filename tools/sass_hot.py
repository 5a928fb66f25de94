"""Per-instruction view of an ncu source-page CSV: executed warp-instructions and stall samples,
grouped into address ranges so phases of a kernel can be compared.
usage: python tools/sass_hot.py x.ncu-rep [n_buckets]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
ia, iS, iN, iE = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
data = [r for r in rows[2:] if len(r) == len(h)]
tot_e = sum(int(r[iE]) for r in data); tot_s = sum(int(r[iN]) for r in data)
print("total warp-instr", tot_e, "samples", tot_s, "static instrs", len(data))
def cls(op):
    op = op.split()[0] if not op.startswith("@") else op.split()[1]
    b = op.split(".")[0]
    if b in ("DADD", "DFMA", "DMUL"): return "DP"
    if b in ("LDS", "STS", "LDG", "STG", "LDSM", "ATOMS"): return "MEM"
    if b in ("SHFL", "VOTE", "REDUX"): return "WARP"
    if b in ("BRA", "BSSY", "BSYNC", "WARPSYNC", "ENDCOLLECTIVE", "BAR", "NOP", "EXIT", "SYNCS"): return "CTRL"
    if b in ("F2F", "I2F", "F2I", "MUFU", "FRND"): return "XU"
    if b in ("FFMA", "FADD", "FMUL", "FFMA2", "FADD2", "FMUL2"): return "FP32"
    return "INT"
agg = {}
for r in data:
    c = cls(r[iS].strip()); agg[c] = agg.get(c, 0) + int(r[iE])
print({k: round(v / tot_e, 3) for k, v in sorted(agg.items(), key=lambda x: -x[1])})
per = (len(data) + nb - 1) // nb
for b in range(0, len(data), per):
    ch = data[b:b + per]
    e = sum(int(r[iE]) for r in ch); s = sum(int(r[iN]) for r in ch)
    dp = sum(int(r[iE]) for r in ch if cls(r[iS].strip()) == "DP")
    print(f"[{b:5d}..{b + len(ch) - 1:5d}] exec {100 * e / tot_e:5.1f}% dp-share {100 * dp / max(e, 1):5.1f}%  samples {100 * s / tot_s:5.1f}%  first: {ch[0][iS].strip()[:50]}")
