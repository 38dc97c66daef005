#!/usr/bin/env python
"""Sweep of the render kernel's scheduler knobs (rr_set_tuning) on one workload; prints one JSON line per setting.

    python tools/tune.py [--workload c4] [--width 1920 --height 1080 --spp 8] [--stats]
"""
import argparse
import itertools
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
ap.add_argument("--spp", type=int, default=8)
ap.add_argument("--bounces", type=int, default=0)
ap.add_argument("--stats", action="store_true")
ap.add_argument("--grid", default="default")
a = ap.parse_args()
kw = dict(width=a.width, height=a.height, spp=a.spp)
if a.bounces:
    kw["bounces"] = a.bounces
wl = workloads.WORKLOADS[a.workload](**kw)
r = rr.Renderer((0,))
r.upload(wl.scene)

PH = ["pixel", "shade", "setup", "trav", "leaf"]


def run(tune, label):
    r.set_tuning(tune)
    r.render_device(wl.cam, wl.width, wl.height, wl.spp, wl.bounces)  # warm
    best = None
    for _ in range(2):
        st = r.render_device(wl.cam, wl.width, wl.height, wl.spp, wl.bounces)
        if best is None or st["render_ms"] < best["render_ms"]:
            best = st
    out = {"tune": list(tune), "label": label, "ms": round(best["render_ms"], 3),
           "mrays_s": round(best["rays"] / best["render_ms"] / 1e3, 1)}
    if a.stats:
        _, _, cs = r.render(wl.cam, wl.width, wl.height, wl.spp, wl.bounces, count_tests=True)
        rays = max(cs["rays"], 1)
        out["box_per_ray"] = round(cs["box_tests"] / rays, 2)
        out["tri_per_ray"] = round(cs["tri_tests"] / rays, 2)
        out["runs_per_ray"] = {n: round(32 * x / rays, 2) for n, x in zip(PH, cs["phase_runs"])}
        out["lanes"] = {n: round(l / max(x, 1), 1) for n, l, x in zip(PH, cs["phase_lanes"], cs["phase_runs"])}
    print(json.dumps(out), flush=True)


#        wP wH wS wT wL keep spec ctas
base = [4, 4, 4, 4, 4, 16, 1, 0]
run(base, "base")
if a.grid == "default":
    for keep in (8, 12, 20, 24, 28, 32):
        t = list(base); t[5] = keep
        run(t, f"keep={keep}")
    for keep in (8, 16, 24, 32):
        t = list(base); t[5] = keep; t[6] = 3
        run(t, f"prefetch keep={keep}")
    t = list(base); t[6] = 0
    run(t, "spec=0")
    for ctas in (3, 4, 5):
        t = list(base); t[7] = ctas
        run(t, f"ctas={ctas}")
    for i, name in ((3, "wT"), (4, "wL"), (2, "wS"), (1, "wH")):
        for w in (2, 8):
            t = list(base); t[i] = w
            run(t, f"{name}={w}")
    run(base, "base again")
elif a.grid == "pixel":  # weight of the pixel phase: how many free slots a warp collects before it refills
    for w in (5, 6, 8, 12, 16, 3):
        t = list(base); t[5] = 20; t[0] = w
        run(t, f"wP={w} keep=20")
    t = list(base); t[5] = 20
    run(t, "base keep=20")
