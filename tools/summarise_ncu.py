"""Turn an .ncu-rep (ncu --set full) into the small text summary kept under profiles/.
usage: python tools/summarise_ncu.py gpurun_out/x.ncu-rep profiles/r1_x.txt ["note"]"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active", "smsp__inst_executed.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__warps_eligible.avg.per_cycle_active",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none summary of {rep}\n# {note}\n")
        for r in rows[2:]:
            f.write(f"\n== {r[hdr.index('Kernel Name')]}\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"{k} = {r[i]} {units[i]}\n")
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        srows = list(csv.reader(io.StringIO(src)))
        if len(srows) > 3 and "# Samples" in srows[1]:
            h = srows[1]
            si, so = h.index("# Samples"), h.index("Source")
            stall = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
            data = [r for r in srows[2:] if len(r) == len(h)]
            tot = sum(int(r[si]) for r in data) or 1
            agg = {}
            for r in data:
                for i in stall:
                    v = int(r[i] or 0)
                    if v:
                        agg[h[i]] = agg.get(h[i], 0) + v
            f.write("\n-- warp stall samples (first kernel), share of all samples\n")
            for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]:
                f.write(f"{k} = {100 * v / tot:.1f} %\n")
            f.write("\n-- hottest SASS instructions (samples)\n")
            for r in sorted(data, key=lambda r: -int(r[si]))[:12]:
                f.write(f"{r[si]:>8}  {r[so].strip()[:90]}\n")
            sass = " ".join(r[so] for r in data)
            f.write("\n-- SASS evidence: " + ", ".join(f"{m}:{sass.count(m)}" for m in ("UBLKCP", "UTMALDG", "SYNCS", "DFMA", "DADD", "FFMA", "HMMA", "UTCHMMA", "REDUX", "ATOMS")) + "\n")


if __name__ == "__main__":
    main()
