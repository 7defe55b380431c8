#!/usr/bin/env python3
"""Per-source-line instruction and stall-sample shares from an .ncu-rep (needs -lineinfo + --import-source on).
usage: tools/ncu_lines.py rep [ntop] [inst|samp]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40; key = sys.argv[3] if len(sys.argv) > 3 else "inst"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(r for r in rows if r and r[0] == "Line No")
ii, si = h.index("Instructions Executed"), h.index("# Samples")
data = [r for r in rows if len(r) > ii and r[0].isdigit() and r[ii].isdigit() and r[si].isdigit()]
tot = sum(int(r[ii]) for r in data); ts = sum(int(r[si]) for r in data)
print(f"warp-instr {tot} samples {ts}")
k = ii if key == "inst" else si
for r in sorted(data, key=lambda r: -int(r[k]))[:ntop]:
    print(f"L{r[0]:>4} inst {100*int(r[ii])/tot:5.2f}% samp {100*int(r[si])/ts:5.2f}% | {r[1].strip()[:110]}")
