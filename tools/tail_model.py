#!/usr/bin/env python
"""The drain of the persistent slot pool, measured on ONE GPU ("prove the drain").

    python tools/tail_model.py [--workload c4] [--spp 256] [--worlds 1,2,4,8]

A pixel's samples are serial (the RNG state is carried from sample to sample, src/Trace.cl:632, 639-642), so the
last pixels a warp took from the queue keep it busy for one pixel lifetime while nothing new arrives.  That tail does
not depend on how many GPUs share the frame, the time before it does -- it is what bends the 1 -> 8 GPU curve.  Here one
GPU renders the tiles t = 0 (mod N) of the frame (rr_render_strided: the same pixel mix as rank 0 of N, no other
rank needed) with the instrumented kernel, which records per warp the time between its first failed tile pop and
its exit.  Prints one JSON line per N: kernel time, the scaling efficiency that time implies (T_1 / (N T_N)), mean and
maximum tail per warp.
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
ap.add_argument("--spp", type=int, default=256)
ap.add_argument("--worlds", default="1,2,4,8")
a = ap.parse_args()
wl = workloads.WORKLOADS[a.workload](spp=a.spp)
r = rr.Renderer((0,))
r.upload(wl.scene)
r.render_strided(wl.cam, wl.width, wl.height, 2, wl.bounces, 0, 1)  # warm-up
base = None
for n in [int(x) for x in a.worlds.split(",")]:
    plain = r.render_strided(wl.cam, wl.width, wl.height, wl.spp, wl.bounces, 0, n)       # production kernel: the time
    r.set_tuning([4, 4, 4, 4, 4, 20, 1 | 8])
    inst = r.render_strided(wl.cam, wl.width, wl.height, wl.spp, wl.bounces, 0, n)        # instrumented kernel: the tail
    r.set_tuning(None)
    if base is None:
        base = plain["render_ms"] * n
    print(json.dumps({"workload": wl.name, "frame": f"{wl.width}x{wl.height}x{wl.spp}spp", "share": f"1/{n} of the tiles",
                      "render_ms": round(plain["render_ms"], 2), "mrays_s": round(plain["rays"] / plain["render_ms"] / 1e3, 1),
                      "implied_efficiency_at_n_gpus": round(base / (n * plain["render_ms"]), 4),
                      "tail_avg_ms": round(inst["tail_avg_ms"], 2), "tail_max_ms": round(inst["tail_max_ms"], 2),
                      "tail_avg_share_of_kernel": round(inst["tail_avg_ms"] / inst["render_ms"], 4),
                      "instrumented_render_ms": round(inst["render_ms"], 2)}), flush=True)
r.close()
