#!/usr/bin/env python
"""Summarise an `ncu --set full` report (exported with `ncu -i X.ncu-rep --page raw --csv`) as a
markdown table: one row per captured launch with duration, DRAM traffic, achieved GB/s and the
limiter-relevant percentages.  Usage: tools/ncu_summary.py raw.csv [bytes_per_pass] > profiles/X.md"""
import csv
import sys

COLS = [("gpu__time_duration.sum", "ms"), ("dram__bytes_read.sum", "GB rd"), ("dram__bytes_write.sum", "GB wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex %"),
        ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "warp-inst")]


def conv(v, unit, want):
    v = float(v.replace(",", ""))
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1., "s": 1e3}
    if want == "ms":
        return v*scale.get(unit, 1.)
    if want.startswith("GB"):
        return v*{"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.}.get(unit, 1e-9)
    return v


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    pass_bytes = float(sys.argv[2]) if len(sys.argv) > 2 else None
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("| kernel | " + " | ".join(c[1] for c in COLS) + " | traffic GB/s |" + (" passes |" if pass_bytes else ""))
    print("|---|" + "---|"*(len(COLS) + 1 + (1 if pass_bytes else 0)))
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].replace("void ", "").replace("mhh::", "").split("(")[0]
        vals = []
        for m, w in COLS:
            vals.append(conv(r[idx[m]], units[idx[m]], w) if m in idx else float("nan"))
        ms, rd, wr = vals[0], vals[1], vals[2]
        cells = [f"{v:.3f}" if i < 3 else (f"{v:.0f}" if v >= 100 else f"{v:.1f}") for i, v in enumerate(vals)]
        line = f"| `{name}` | " + " | ".join(cells) + f" | {(rd+wr)/(ms*1e-3):.0f} |"
        if pass_bytes:
            line += f" {(rd+wr)*1e9/pass_bytes:.2f} |"
        print(line)


if __name__ == "__main__":
    main()
