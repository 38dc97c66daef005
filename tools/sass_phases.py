#!/usr/bin/env python
"""Static code-size map of k_render: SASS instructions per scheduler phase, from `nvdisasm -g -c` line info.

    python tools/sass_phases.py [path/to/librr_b200.so]
The SM instruction cache holds 32 KB (2 048 instructions); the phases a warp cycles through must fit.
"""
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
so = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "ripoff_raytracer_b200/csrc/librr_b200.so"
src = (ROOT / "ripoff_raytracer_b200/csrc/rr_render.cu").read_text().splitlines()


def line_of(pat):
    for i, l in enumerate(src, 1):
        if pat in l:
            return i
    raise SystemExit(f"anchor not found: {pat}")


anchors = [("vote", line_of("---- vote:")), ("trav", line_of("================= node steps")),
           ("leaf", line_of("================= leaf tests")), ("setup", line_of("================= finish the current mesh")),
           ("shade", line_of("================= one bounce of Trace()")), ("pixel", line_of("================= hand a pixel")),
           ("end", line_of("#undef CW"))]
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "rr_render.sm_100a.cubin", str(so)], cwd=td, check=True, capture_output=True)
    cub = next(Path(td).glob("*.cubin"))
    out = subprocess.run(["nvdisasm", "-g", "-c", str(cub)], capture_output=True, text=True).stdout
fn = None
cur_line = None
cur_file = None
counts = {}
phase = "prolog"
for l in out.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+)", l)
    if m:
        fn = m.group(1).split(",")[0]
        phase = "prolog"
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur_file, cur_line = m.group(1).split("/")[-1], int(m.group(2))
        if cur_file == "rr_render.cu":
            for (name, lo), (_, hi) in zip(anchors, anchors[1:]):
                if lo <= cur_line < hi:
                    phase = name
        continue
    if re.match(r"\s+/\*[0-9a-f]{4}\*/", l) and fn:
        counts.setdefault(fn, {}).setdefault(phase, 0)
        counts[fn][phase] += 1
for fn, c in counts.items():
    if "k_render" not in fn and sum(c.values()) < 40:
        continue
    tot = sum(c.values())
    print(f"{fn[:70]:70} {tot:5d} instr {tot*16/1024:6.1f} KB  " + " ".join(f"{k}={v}" for k, v in c.items()))
