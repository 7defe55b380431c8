#!/usr/bin/env python3
"""Summarise an .ncu-rep: key raw metrics + top SASS lines by stall samples / instruction share.
usage: tools/ncu_summary.py gpurun_out/prof_x.ncu-rep [ntop]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"]
print("kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
for i, h in enumerate(hdr):
    if h in want or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
        try:
            if "issue_stalled" in h and float(vals[i]) < 0.05: continue
        except ValueError: pass
        print(f"  {h:86s} {units[i]:14s} {vals[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
r = csv.reader(io.StringIO(src)); next(r); h = next(r)
rows = [x for x in r if len(x) > 6]
tot = sum(int(x[5]) for x in rows); ts = sum(int(x[4]) for x in rows)
print(f"SASS lines {len(rows)}  warp-instr {tot}  samples {ts}")
idx = sorted(range(len(rows)), key=lambda i: -int(rows[i][4]))[:ntop]
for i in idx:
    x = rows[i]
    print(f"  {i:5d} samp {100*int(x[4])/ts:5.2f}% inst {100*int(x[5])/tot:5.2f}%  {x[1].strip()}")
if len(sys.argv) > 3:
    with open(sys.argv[3], "w") as f:
        for i, x in enumerate(rows):
            f.write(f"{i:5d} {100*int(x[5])/tot:5.2f}% s{100*int(x[4])/ts:5.2f}% {x[1].strip()}\n")
