#!/usr/bin/env python
"""One short invocation of the hot path for ncu: upload a workload, render N frames, print the counters.

    python tools/profile_run.py [--workload c4] [--width 1920 --height 1080 --spp 4] [--frames 1] [--count]
"""
import argparse
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import ripoff_raytracer_b200 as rr  # noqa: E402
from ripoff_raytracer_b200 import workloads  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c4")
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--spp", type=int, default=4)
ap.add_argument("--bounces", type=int, default=0)
ap.add_argument("--frames", type=int, default=1)
ap.add_argument("--count", action="store_true")
ap.add_argument("--generic", action="store_true", help="force the kernel instantiation with every scene feature compiled in (tuning bit 6)")
a = ap.parse_args()
kw = dict(width=a.width, height=a.height, spp=a.spp)
if a.bounces:
    kw["bounces"] = a.bounces
wl = workloads.WORKLOADS[a.workload](**kw)
r = rr.Renderer((0,))
r.upload(wl.scene)
if a.generic:
    r.set_tuning([4, 4, 4, 4, 4, 20, 1 | 64])
for _ in range(a.frames):
    if a.count:
        _, _, st = r.render(wl.cam, wl.width, wl.height, wl.spp, wl.bounces, count_tests=True)
    else:
        st = r.render_device(wl.cam, wl.width, wl.height, wl.spp, wl.bounces)
    rays = st["rays"]
    st["mrays_s"] = rays / st["render_ms"] / 1e3
    st["msamples_s"] = st["samples"] / st["render_ms"] / 1e3
    print(json.dumps(st))
