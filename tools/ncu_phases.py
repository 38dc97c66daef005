#!/usr/bin/env python
"""Attribute the SASS of k_render to scheduler phases from an
`ncu --page source --csv --print-source cuda,sass` export: instructions are walked in address order and
labelled by the last kernel-body source line seen (inlined helpers inherit the phase of their call site).

    python tools/ncu_phases.py gpurun_out/src.csv vote:345-367 trav:368-392 ...
"""
import csv
import sys

path = sys.argv[1]
ranges = []
for spec in sys.argv[2:]:
    name, r = spec.split(":")
    lo, hi = r.split("-")
    ranges.append((name, int(lo), int(hi)))
rows = list(csv.reader(open(path)))
hdr = None
cur_file, cur_line = None, None
data = []


def num(x):
    try:
        return int(float(x))
    except ValueError:
        return 0


for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        iA, iI, iT, iS = 2, hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
        continue
    if r[0] == "Function Name" or hdr is None:
        continue
    if r[0] != "":
        try:
            cur_line = int(r[0])
        except ValueError:
            pass
        continue
    if r[iA].startswith("0x"):
        data.append((int(r[iA], 16), cur_file, cur_line, num(r[iI]), num(r[iT]), num(r[iS]), r[3]))
data.sort()


def phase_of(f, line):
    if f != "rr_render.cu":
        return None
    for name, lo, hi in ranges:
        if lo <= line <= hi:
            return name
    return None


tot = {}
cur = "prolog"
for addr, f, line, wi, ti, s, sass in data:
    p = phase_of(f, line)
    if p and not p.startswith("_"):
        cur = p
    key = cur + ("/" + p[1:] if p and p.startswith("_") else "")
    t = tot.setdefault(key, [0, 0, 0, 0])
    t[0] += wi; t[1] += ti; t[2] += s; t[3] += 1
W = sum(t[0] for t in tot.values()) or 1
S = sum(t[2] for t in tot.values()) or 1
print(f"total warp inst {W:,}   samples {S:,}")
for k, t in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:18} sass={t[3]:5d} winst%={100*t[0]/W:6.2f} thr/inst={t[1]/max(t[0],1):5.1f} samp%={100*t[2]/S:6.2f}")
