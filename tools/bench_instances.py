#!/usr/bin/env python
"""cost of many instances -- top level (Morton chunks + chunk / group boxes) against the linear
mesh scan of the same kernel (the reference's loop over meshCount, src/Trace.cl:444).  One JSON line per setting.

    python tools/bench_instances.py [--counts 64,1024,8192] [--width 1920 --height 1080 --spp 4]
"""
import argparse
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import ripoff_raytracer_b200 as rr  # noqa: E402
from ripoff_raytracer_b200 import workloads  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--counts", default="64,1024,8192")
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--spp", type=int, default=4)
ap.add_argument("--bounces", type=int, default=16)
ap.add_argument("--linear-max", type=int, default=2048, help="largest count the linear scan is timed on")
a = ap.parse_args()

r = rr.Renderer((0,))
for count in [int(c) for c in a.counts.split(",")]:
    wl = workloads.instances(width=a.width, height=a.height, spp=a.spp, bounces=a.bounces, count=count)
    t0 = time.perf_counter()
    r.upload(wl.scene)
    upload_ms = (time.perf_counter() - t0) * 1e3
    meshes = wl.scene.arrays()[1]
    t0 = time.perf_counter()
    r.update_meshes(meshes)
    update_ms = (time.perf_counter() - t0) * 1e3
    for label, tune in (("top_level", None), ("linear", [4, 4, 4, 4, 4, 20, 1 | 4])):
        if label == "linear" and count > a.linear_max:
            continue
        r.set_tuning(tune)
        r.render_device(wl.cam, wl.width, wl.height, 1, wl.bounces)  # warm
        st = min((r.render_device(wl.cam, wl.width, wl.height, wl.spp, wl.bounces) for _ in range(2)), key=lambda s: s["render_ms"])
        _, _, cs = r.render(wl.cam, wl.width, wl.height, 1, wl.bounces, count_tests=True)
        print(json.dumps({"instances": count, "meshes": int(len(meshes)), "mode": label, "frame": f"{a.width}x{a.height}x{a.spp}spp",
                          "render_ms": round(st["render_ms"], 3), "mrays_s": round(st["rays"] / st["render_ms"] / 1e3, 1),
                          "box_per_ray": round(cs["box_tests"] / max(cs["rays"], 1), 1),
                          "tri_per_ray": round(cs["tri_tests"] / max(cs["rays"], 1), 2),
                          "upload_ms": round(upload_ms, 2), "update_meshes_ms": round(update_ms, 3)}), flush=True)
    r.set_tuning(None)
r.close()
