#!/usr/bin/env python
"""Condenses `ncu -i X.ncu-rep --page raw --csv` (and, optionally, `--page source --csv`) of ONE kernel launch into the
handful of numbers DESIGN.md / bench.py quote: duration, DRAM traffic, registers, occupancy, pipe utilisation, shared-memory
wavefronts and bank conflicts, opcode mix.   usage: ncu_summary.py raw.csv [source.csv] > profiles/rNN/ncu_<kernel>.txt"""
import collections
import csv
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "LSU data-pipe wavefronts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput"),
    ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue slots busy"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma pipe"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu pipe"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block", "shared memory / block"),
    ("launch__grid_size", "grid size"),
    ("launch__block_size", "block size"),
    ("launch__occupancy_limit_registers", "blocks/SM limit: registers"),
    ("launch__occupancy_limit_shared_mem", "blocks/SM limit: shared memory"),
    ("launch__occupancy_limit_warps", "blocks/SM limit: warps"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("smsp__inst_executed_op_local_ld.sum", "local-memory loads (spills)"),
    ("smsp__inst_executed_op_local_st.sum", "local-memory stores (spills)"),
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, vals = rows[0], rows[1], rows[2]
    ix = {h: i for i, h in enumerate(hdr)}
    name = vals[ix["Kernel Name"]] if "Kernel Name" in ix else "?"
    print(f"kernel: {name}")
    print(f"source: {sys.argv[1]}  (ncu --set full --clock-control none, one launch; cold-cache, serialised)")
    for key, label in WANT:
        if key in ix:
            print(f"  {label:36s} {vals[ix[key]]:>18s} {units[ix[key]]}")
    if len(sys.argv) > 2:
        src = list(csv.reader(open(sys.argv[2])))
        h = src[1]; jx = {x: i for i, x in enumerate(h)}
        ops = collections.Counter(); tot = 0
        sh = collections.defaultdict(lambda: [0, 0, 0])
        stall = collections.Counter()
        for r in src[2:]:
            if len(r) < len(h):
                continue
            t = r[jx["Source"]].split()
            if not t:
                continue
            op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
            ex = int(r[jx["Instructions Executed"]] or 0)
            ops[op.split(".")[0]] += ex; tot += ex
            if op.startswith("LDS") or op.startswith("STS"):
                a = sh[op]
                a[0] += ex; a[1] += int(r[jx["L1 Wavefronts Shared"]] or 0); a[2] += int(r[jx["L1 Wavefronts Shared Ideal"]] or 0)
            for c in h:
                if c.startswith("stall_"):
                    stall[c] += int(r[jx[c]] or 0)
        print(f"  opcode mix (warp instructions, {tot} total): " + ", ".join(f"{k} {v/tot:.1%}" for k, v in ops.most_common(12)))
        for op, (ex, wf, ideal) in sorted(sh.items(), key=lambda kv: -kv[1][1]):
            print(f"  {op:10s} executed {ex:>12d}  wavefronts {wf:>12d}  ideal {ideal:>12d}  ({wf/max(ideal,1):.2f}x)")
        ts = sum(stall.values()) or 1
        print("  stall samples: " + ", ".join(f"{k[6:]} {v/ts:.1%}" for k, v in stall.most_common(8)))


if __name__ == "__main__":
    main()
