#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics + per-CUDA-source-line stall samples / instructions.
usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep [source_file_for_line_text] [top_n]"""
import csv, subprocess, sys, io

rep = sys.argv[1]
srcfile = sys.argv[2] if len(sys.argv) > 2 else None
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__warps_eligible.avg.per_cycle_active", "launch__shared_mem_per_block_allocated",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_srcunit_tex_op_read.sum"]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print("== kernel:", name[:100])
    for h, u, v in zip(hdr, units, r):
        if h in WANT or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
            print(f"  {h:78s} {u:16s} {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
H = len(hdr)
iS = hdr.index("# Samples") - H
iI = hdr.index("Instructions Executed") - H
agg = {}
for r in rows[hi + 1:]:
    if len(r) < H or not r[0].isdigit():
        continue
    try:
        s, ins = int(r[iS]), int(r[iI])
    except ValueError:
        continue
    a = agg.setdefault(int(r[0]), [0, 0])
    a[0] = max(a[0], s)
    a[1] = max(a[1], ins)
tot = sum(a[0] for a in agg.values()) or 1
toti = sum(a[1] for a in agg.values()) or 1
lines = open(srcfile).read().split("\n") if srcfile else []
print(f"total samples {tot}  total warp-instructions {toti}")
for k, a in sorted(agg.items(), key=lambda x: -x[1][0])[:topn]:
    text = lines[k - 1].strip()[:95] if 0 < k <= len(lines) else ""
    print(f"L{k:4d} samples {100*a[0]/tot:5.1f}%  inst {100*a[1]/toti:5.1f}% ({a[1]:>11d}) | {text}")
