#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into the few numbers DESIGN.md / bench.py quote: duration, DRAM bytes and
throughput, achieved occupancy, issue / pipe utilisation, top stall reasons.  Usage: tools/ncu_summary.py rep [out.txt]
Needs the `ncu` CLI (present in the build image; no GPU required to read a report)."""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__inst_executed.sum", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__cycles_active.avg",
]
STALL = "smsp__average_warp"


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        print("no data in", rep)
        return
    hdr, units = rows[0], rows[1]
    lines = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        lines.append(f"== {d.get('Kernel Name', '?')}  grid {d.get('Grid Size', '?')} block {d.get('Block Size', '?')}")
        for k in KEYS:
            if k in d and d[k] != "":
                lines.append(f"  {k:75s} {d[k]:>18s} {units[hdr.index(k)]}")
        stalls = [(k, d[k]) for k in hdr if k.startswith(STALL) and "ratio" in k and d.get(k, "") not in ("", "0")]
        def fnum(x):
            try:
                return float(x.replace(",", ""))
            except ValueError:
                return 0.0
        stalls.sort(key=lambda kv: -fnum(kv[1]))
        for k, v in stalls[:8]:
            lines.append(f"  {k:75s} {v:>18s}")
        try:
            t = fnum(d["gpu__time_duration.sum"]); rd = fnum(d["dram__bytes_read.sum"]); wr = fnum(d["dram__bytes_write.sum"])
            ut, ub = units[hdr.index("gpu__time_duration.sum")], units[hdr.index("dram__bytes_read.sum")]
            lines.append(f"  -> duration {t} {ut}; DRAM read+write {rd + wr} {ub}")
        except Exception:
            pass
    text = "\n".join(lines)
    print(text)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(f"# {rep} (ncu --set full --clock-control none)\n" + text + "\n")


if __name__ == "__main__":
    main()
