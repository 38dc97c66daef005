#!/usr/bin/env python
"""OBJ text ingest (reference src/readobj.hpp:270-376): parse time of the host loader by thread
count, then the whole path file -> device scene (rr_obj_load + rr_upload_scene_indexed).  One JSON line per setting.

    python tools/bench_obj.py [--grid 500] [--threads 1,2,4,8,16]
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import ripoff_raytracer_b200 as rr  # noqa: E402
from ripoff_raytracer_b200 import _abi, scenes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=500, help="height-field resolution: 2*grid^2 triangles")
ap.add_argument("--threads", default="1,2,4,8,16")
ap.add_argument("--child", type=int, default=0)
ap.add_argument("--obj", default="")
a = ap.parse_args()

if a.child:  # RR_OBJ_THREADS is read per call, but a fresh process keeps the page cache the only shared state
    os.environ["RR_OBJ_THREADS"] = str(a.child)
    best_i = best_s = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        pos, nrm, cor = rr.load_obj_indexed(a.obj)
        best_i = min(best_i, time.perf_counter() - t0)
        s = rr.Scene()
        t0 = time.perf_counter()
        s.load_obj(a.obj)
        best_s = min(best_s, time.perf_counter() - t0)
        s.close()
    mb = os.path.getsize(a.obj) / 1e6
    print(json.dumps({"threads": a.child, "file_mb": round(mb, 1), "triangles": int(len(cor)),
                      "rr_obj_load_ms": round(best_i * 1e3, 1), "rr_obj_load_mb_s": round(mb / best_i, 1),
                      "rr_scene_load_obj_ms": round(best_s * 1e3, 1)}), flush=True)
    sys.exit(0)

with tempfile.TemporaryDirectory() as td:
    obj = Path(td) / "terrain.obj"
    v, n, f = scenes.heightfield(a.grid, size=800.0, height=70.0, base=0.0, seed=4)
    scenes.write_obj(obj, v, n, f)
    for th in [int(x) for x in a.threads.split(",")]:
        subprocess.run([sys.executable, __file__, "--child", str(th), "--obj", str(obj)], check=True)
    # the same call from C (no Python-side copies of the result), with the loader's own phase timing on stderr
    exe = Path(td) / "bench_obj_c"
    lib = ROOT / "ripoff_raytracer_b200" / "csrc"
    subprocess.run(["g++", "-O2", "-o", str(exe), str(ROOT / "tools" / "bench_obj_c.cpp"), f"-L{lib}", "-lrr_b200", f"-Wl,-rpath,{lib}"], check=True)
    for th in [int(x) for x in a.threads.split(",")]:
        out = subprocess.run([str(exe), str(obj)], env=dict(os.environ, RR_OBJ_THREADS=str(th), RR_OBJ_DEBUG="1"), capture_output=True, text=True)
        best = min(float(l.split()[-2]) for l in out.stdout.splitlines() if l.startswith("rc=0"))
        laps = [l for l in out.stderr.splitlines() if l.startswith("[obj]")][-5:]
        print(json.dumps({"threads": th, "c_api_rr_obj_load_ms": round(best * 1e3, 1), "mb_s": round(os.path.getsize(obj) / 1e6 / best, 1),
                          "phases_last_run": "; ".join(l[6:] for l in laps)}), flush=True)
    # file -> device: parse + indexed upload + LBVH build
    try:
        ren = rr.Renderer()
    except _abi.RRError:
        sys.exit(0)  # no GPU here: parse timings only
    m = np.zeros(1, _abi.MESH)
    m["scale"] = 1.0
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        pos, nrm, cor = rr.load_obj_indexed(obj)
        r = np.zeros(1, _abi.MESH_RANGE)
        r["numTriangles"] = len(cor)
        ren.upload_indexed(pos, nrm, cor, m, r)
        best = min(best, time.perf_counter() - t0)
    print(json.dumps({"file_to_device_ms": round(best * 1e3, 1), "triangles": int(len(cor)), "threads": os.cpu_count()}), flush=True)
    ren.close()
