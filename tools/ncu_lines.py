"""Per source line of a kernel in an .ncu-rep (ncu --set full --import-source on, built with -lineinfo): warp-stall samples,
executed warp instructions and shared-memory wavefronts, to see which part of a kernel the time goes to.
usage: python tools/ncu_lines.py x.ncu-rep [top]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    cur_file, hdr, out = None, None, []
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[0] not in ("", "Line No"):
            g = lambda k: r[hdr.index(k)]
            out.append((cur_file, int(r[0]), r[1].strip()[:100], int(g("# Samples") or 0), int(g("Instructions Executed") or 0),
                        int(g("L1 Wavefronts Shared") or 0)))
    tot_s = sum(o[3] for o in out) or 1
    tot_i = sum(o[4] for o in out) or 1
    tot_w = sum(o[5] for o in out) or 1
    print(f"samples {tot_s}  warp instructions {tot_i}  shared wavefronts {tot_w}")
    print("file:line  samples%  inst%  wavefronts%  source")
    for o in sorted(out, key=lambda o: (o[0], o[1])):
        if o[3] * 1000 >= tot_s * 2 or o[5] * 100 >= tot_w:
            print(f"{o[0]}:{o[1]:<5} {100*o[3]/tot_s:6.2f} {100*o[4]/tot_i:6.2f} {100*o[5]/tot_w:6.2f}  {o[2]}")


if __name__ == "__main__":
    main()
