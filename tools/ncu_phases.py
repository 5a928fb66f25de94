"""Summarise an ncu source-page CSV by barrier-delimited phases and top instructions.
usage: python tools/ncu_phases.py file.csv [n_top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 12
si = hdr.index('# Samples'); src = hdr.index('Source'); ie = hdr.index('Instructions Executed')
stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[si]) for r in data)
cur = [0, 0, {}, 0]; n = 0
def flush(label):
    global cur
    if cur[0] * 200 > tot:
        top = sorted(cur[2].items(), key=lambda x: -x[1])[:5]
        print(f"phase {label}: {100*cur[0]/tot:.1f}% samples, {cur[1]/1e6:.1f}M warp-insts, {cur[3]} sass lines; " + ", ".join(f"{k[6:]} {100*v/cur[0]:.0f}%" for k, v in top))
    cur = [0, 0, {}, 0]
for r in data:
    cur[0] += int(r[si]); cur[1] += int(r[ie]); cur[3] += 1
    for i in stall:
        v = int(r[i]) if r[i] else 0
        if v: cur[2][hdr[i]] = cur[2].get(hdr[i], 0) + v
    if 'BAR.' in r[src] or 'EXIT' in r[src] or 'SYNCS' in r[src] or 'WARPSYNC' in r[src]:
        flush(f"{n} ..{r[src].strip()[:44]}"); n += 1
flush('last')
print('total samples', tot, 'total warp-insts %.1fM' % (sum(int(r[ie]) for r in data) / 1e6))
for r in sorted(data, key=lambda r: -int(r[si]))[:ntop]:
    print(r[si].rjust(8), r[src][:90])
