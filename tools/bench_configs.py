#!/usr/bin/env python
"""Kernel-only throughput of every BASELINE.json configuration at a reduced frame (one JSON line each).

    python tools/bench_configs.py [--spp 8] [--width 1920 --height 1080] [c1 c2 c3 c4 c5]
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
ap.add_argument("names", nargs="*", default=["c1", "c2", "c3", "c4", "c5"])
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--spp", type=int, default=8)
a = ap.parse_args()
for name in a.names:
    t0 = time.time()
    wl = workloads.WORKLOADS[name](width=a.width, height=a.height, spp=a.spp)
    gen_s = time.time() - t0
    r = rr.Renderer((0,))
    t0 = time.time()
    r.upload(wl.scene)
    up_s = time.time() - t0
    r.render_device(wl.cam, wl.width, wl.height, wl.spp, wl.bounces)
    st = min((r.render_device(wl.cam, wl.width, wl.height, wl.spp, wl.bounces) for _ in range(2)), key=lambda s: s["render_ms"])
    img, _, cs = r.render(wl.cam, wl.width, wl.height, min(wl.spp, 2), wl.bounces, count_tests=True)
    rays = max(cs["rays"], 1)
    print(json.dumps({"workload": wl.name, "triangles": wl.scene.n_triangles, "spheres": wl.scene.n_spheres,
                      "frame": f"{wl.width}x{wl.height}x{wl.spp}spp", "mrays_s": round(st["rays"] / st["render_ms"] / 1e3, 1),
                      "msamples_s": round(st["samples"] / st["render_ms"] / 1e3, 1), "render_ms": round(st["render_ms"], 2),
                      "rays_per_sample": round(st["rays"] / st["samples"], 3), "build_ms": round(st["build_ms"], 2),
                      "upload_s": round(up_s, 3), "generate_s": round(gen_s, 2),
                      "box_per_ray": round(cs["box_tests"] / rays, 2), "tri_per_ray": round(cs["tri_tests"] / rays, 2),
                      "sphere_per_ray": round(cs["sphere_tests"] / rays, 2),
                      "lit_pixels": float((img[..., :3].max(-1) > 0).mean())}), flush=True)
    r.close()
