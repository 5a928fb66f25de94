"""SASS opcode histogram of every kernel in bliss_b200/libbliss.so (cuobjdump, CPU box): what the evidence table of
/opt/skills/guides/B200_PROFILING.md asks for - UTMALDG / UBLKCP (TMA), SYNCS (mbarrier), FADD2, DFMA ... per kernel.
usage: python tools/sass_histogram.py > profiles/r2_sass_ops.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "bliss_b200", "libbliss.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern).replace("blx::", "").replace("(anonymous namespace)::", "")
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_]+)*)", line)
    if m and kern:
        hist[kern][m.group(1).split(".")[0]] += 1
print(f"# SASS opcode histogram (static instruction counts) of {os.path.relpath(so, ROOT)}, sm_100a, cuobjdump -sass")
mark = ("UTMALDG", "UBLKCP", "SYNCS", "FADD2", "FFMA2", "DFMA", "DADD", "DMUL", "ATOMS", "RED", "REDUX", "STG", "LDG", "LDS", "STS", "SHFL",
        "F2F", "I2F", "F2I", "FRND", "IMAD", "HMMA", "UTCHMMA")
for k, c in hist.items():
    tot = sum(c.values())
    print(f"\n== {k}: {tot} instructions")
    print("   " + ", ".join(f"{m}:{c[m]}" for m in mark if c[m]))
    print("   top: " + ", ".join(f"{o}:{n}" for o, n in c.most_common(12)))
