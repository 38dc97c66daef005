#!/usr/bin/env python
"""Driver of tools/bvh_quality_experiment.cpp: dumps the triangles of a BASELINE scene, compiles and runs the experiment
(CPU only, ~15 s).  Result of round 2: profiles/bvh_quality_r02_c4.txt."""
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from ripoff_raytracer_b200 import workloads  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c4"
wl = workloads.WORKLOADS[name](width=480, height=270, spp=1)
t, m, r, _ = wl.scene.arrays()
big = int(np.argmax(r["numTriangles"]))
f, c = int(r["firstTriangle"][big]), int(r["numTriangles"][big])
P = np.stack([t["posA"][:, :3], t["posB"][:, :3], t["posC"][:, :3]], 1).astype(np.float32)
with tempfile.TemporaryDirectory() as td:
    P[f:f + c].tofile(f"{td}/{name}_big.bin")
    np.concatenate([P[:f], P[f + c:]]).tofile(f"{td}/{name}_other.bin")
    subprocess.run(["g++", "-O2", "-std=c++17", "-w", "-o", f"{td}/exp", str(ROOT / "tools" / "bvh_quality_experiment.cpp")], check=True)
    subprocess.run(f"ulimit -s unlimited; ./exp {name}", shell=True, cwd=td, check=True)
