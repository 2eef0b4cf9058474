"""Summarise an .ncu-rep: headline metrics, stall-reason shares, instruction mix and the top stalled SASS lines.
    python profiles/ncu_summary.py report.ncu-rep [units_per_kernel]   (units = warp-steps to normalise the instruction mix)"""
import collections
import csv
import io
import subprocess
import sys


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    units = float(sys.argv[2]) if len(sys.argv) > 2 else None
    det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
    keys = ("Duration", "DRAM Throughput", "Executed Ipc Active", "Issue Slots Busy", "No Eligible", "Active Warps Per Scheduler", "Registers Per Thread",
            "Achieved Occupancy", "Warp Cycles Per Issued", "Mem Busy", "L1/TEX Hit Rate", "L2 Hit Rate", "Max Bandwidth")
    for line in det.splitlines():
        if any(k in line for k in keys) and "OPT" not in line and "INF" not in line:
            print(" ".join(line.split()))
    raw = page(rep, "raw")
    if len(raw) > 2:
        d = dict(zip(raw[0], raw[2]))
        for k in ("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
                  "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
                  "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
                  "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum"):
            if k in d:
                print(k, d[k])
    rows = page(rep, "source")
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def f(r, k):
        try:
            return float(r[ix[k]])
        except Exception:
            return 0.0

    tot = sum(f(r, "# Samples") for r in data) or 1.0
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {s: sum(f(r, s) for r in data) for s in stalls}
    print("stalls:", ", ".join(f"{s[6:]} {100 * v / tot:.1f}%" for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]))
    ti = sum(f(r, "Instructions Executed") for r in data)
    ops = collections.Counter()
    for r in data:
        src = r[ix["Source"]].split()
        op = [w for w in src if not w.startswith("@")][0] if src else "?"
        ops[op.split(".")[0]] += f(r, "Instructions Executed")
    div = units or 1.0
    print(f"warp instructions {ti:.0f}" + (f" = {ti / div:.1f} per unit" if units else ""))
    print("mix:", ", ".join(f"{k} {v / div:.1f}" if units else f"{k} {100 * v / ti:.1f}%" for k, v in ops.most_common(18)))
    for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:14]:
        top = max(stalls, key=lambda s: f(r, s))
        print(f"{100 * f(r, '# Samples') / tot:5.1f}% {top[6:]:14s} {r[ix['Source']][:100]}")


if __name__ == "__main__":
    main()
