#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line.

    python tools/ncu_lines.py gpurun_out/src.csv [top]
Prints warp instructions, thread instructions and stall samples per source line of rr_render.cu (and inlined headers).
"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
# the export is a sequence of blocks: "File Path",<file> / "Function Name",<fn> / header / lines...
agg = defaultdict(lambda: [0, 0, 0, ""])
cur_file = None
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        hdr = None
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        i_src = 1
        i_samp = hdr.index("# Samples")
        i_inst = hdr.index("Instructions Executed")
        i_thr = hdr.index("Thread Instructions Executed")
        continue
    if hdr is None:
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    def num(x):
        try:
            return int(float(x))
        except ValueError:
            return 0
    key = (cur_file, line)
    a = agg[key]
    a[0] += num(r[i_inst]); a[1] += num(r[i_thr]); a[2] += num(r[i_samp]); a[3] = r[i_src].strip()[:110]
tot_i = sum(a[0] for a in agg.values()) or 1
tot_s = sum(a[2] for a in agg.values()) or 1
print(f"total warp inst {tot_i:,}  samples {tot_s:,}")
print(f"{'file:line':28} {'winst%':>7} {'thr/inst':>8} {'samp%':>6}  source")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    print(f"{key[0]+':'+str(key[1]):28} {100*a[0]/tot_i:7.2f} {a[1]/max(a[0],1):8.1f} {100*a[2]/tot_s:6.2f}  {a[3]}")
