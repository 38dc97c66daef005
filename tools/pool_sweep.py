#!/usr/bin/env python
"""Slots per warp in use vs the share of the frame one GPU renders (wave quantisation of the slot pool).

    python tools/pool_sweep.py [--workload c4] [--spp 64] [--worlds 8,4,2,1] [--pools 96,92,88,84,80,76,72]

A pixel occupies a path slot for its whole lifetime (the samples of a pixel are serial), every slot starts its first
pixel at t = 0, and pixel lifetimes differ little: the slots of a GPU work through the frame in waves.  When the pixels
of a GPU's share are not a whole number of waves, the last wave runs with part of the pool idle.  One GPU renders the
tiles t = 0 (mod N) of the frame (rr_render_strided, the share of one of N ranks) with 96 ... 72 of the 96 slots per
warp taking pixels (rr_set_tuning value 8); one JSON line per (N, slots): kernel time and the waves that many slots make.
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
ap.add_argument("--spp", type=int, default=64)
ap.add_argument("--worlds", default="8,4,2,1")
ap.add_argument("--pools", default="96,92,88,84,80,76,72")
ap.add_argument("--auto", action="store_true", help="also time pool 0 = the library's own choice")
ap.add_argument("--repeat", type=int, default=2)
a = ap.parse_args()
wl = workloads.WORKLOADS[a.workload](spp=a.spp)
r = rr.Renderer((0,))
r.upload(wl.scene)
r.render_strided(wl.cam, wl.width, wl.height, 2, wl.bounces, 0, 1)  # warm-up
warps = 148 * 20
for n in [int(x) for x in a.worlds.split(",")]:
    pixels = wl.width * wl.height / n
    pools = [int(x) for x in a.pools.split(",")] + ([0] if a.auto else [])
    for pool in pools:
        r.set_tuning([4, 4, 4, 4, 4, 20, 1, 0, pool])
        ms = [r.render_strided(wl.cam, wl.width, wl.height, wl.spp, wl.bounces, 0, n) for _ in range(a.repeat)]
        best = min(ms, key=lambda s: s["render_ms"])
        print(json.dumps({"workload": wl.name, "frame": f"{wl.width}x{wl.height}x{wl.spp}spp", "share": f"1/{n}", "pool_use": pool,
                          "waves": round(pixels / (warps * pool), 3) if pool else None,
                          "render_ms": [round(s["render_ms"], 2) for s in ms],
                          "mrays_s": round(best["rays"] / best["render_ms"] / 1e3, 1)}), flush=True)
r.set_tuning(None)
r.close()
