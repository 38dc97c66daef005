#!/usr/bin/env python
"""What would a cost-ordered queue buy?  (heaviest tiles first.)

    python tools/cost_order.py [--workload c4] [--spp 64] [--pre-spp 4] [--block 16] [--worlds 1,8]

A pre-pass renders ONE pixel per tile (a frame of tiles_x x tiles_y pixels, same camera: pixel (i, j) looks along the first
pixel of tile (i, j)) with the instrumented kernel and reads back the path segments per pixel (rr_render_cost).  Tiles are
grouped into blocks of B x B tiles, the blocks sorted by mean cost, and the frame (or the 1/N share of it that rank 0 of N
would render: tiles t = 0 mod N) is rendered with the queue handing out the tiles in that order (rr_set_tile_order):
heaviest block first, lightest first (control), and the default row-major order.  Prints one JSON line per case with the
kernel time and the drain per warp (instrumented kernel).
"""
import argparse
import json
import sys
import time
import zlib
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import ripoff_raytracer_b200 as rr  # noqa: E402
from ripoff_raytracer_b200 import multigpu, workloads  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c4")
ap.add_argument("--spp", type=int, default=64)
ap.add_argument("--pre-spp", type=int, default=4)
ap.add_argument("--block", type=int, default=16)
ap.add_argument("--worlds", default="1,8")
a = ap.parse_args()
wl = workloads.WORKLOADS[a.workload](spp=a.spp)
W, H = wl.width, wl.height
tx, ty = multigpu.tile_grid(W, H)
r = rr.Renderer((0,))
r.upload(wl.scene)
r.render_device(wl.cam, W, H, 2, wl.bounces)  # warm-up
r.render_cost(wl.cam, tx, ty, 1, wl.bounces)
t0 = time.perf_counter()
cost = r.render_cost(wl.cam, tx, ty, a.pre_spp, wl.bounces).astype(np.float64)  # (ty, tx): cost of tile j * tx + i
pre_ms = (time.perf_counter() - t0) * 1e3
B = a.block
by, bx = -(-ty // B), -(-tx // B)
blocks = []
for j in range(by):
    for i in range(bx):
        c = cost[j * B:(j + 1) * B, i * B:(i + 1) * B]
        tiles = [(j * B + v) * tx + (i * B + u) for v in range(c.shape[0]) for u in range(c.shape[1])]
        blocks.append((float(c.mean()), tiles))
print(json.dumps({"pre_pass_ms_wall": round(pre_ms, 2), "pre_pass": f"{tx}x{ty} pixels x {a.pre_spp} spp", "blocks": len(blocks),
                  "block_cost_min_mean_max": [round(min(b[0] for b in blocks) / a.pre_spp, 2), round(float(cost.mean()) / a.pre_spp, 2),
                                              round(max(b[0] for b in blocks) / a.pre_spp, 2)],
                  "row_cost_top_to_bottom": [round(float(x) / a.pre_spp, 2) for x in cost.reshape(9, -1, tx).mean(axis=(1, 2))] if ty % 9 == 0 else None}), flush=True)
orders = {"row-major": None,
          "heaviest block first": [t for _, ts in sorted(blocks, key=lambda b: -b[0]) for t in ts],
          "lightest block first": [t for _, ts in sorted(blocks, key=lambda b: b[0]) for t in ts]}
crc0 = None
for n in [int(x) for x in a.worlds.split(",")]:
    for name, order in orders.items():
        table = None if order is None else np.array([t for t in order if t % n == 0], np.uint32)
        r.set_tile_order(table)
        plain = min((r.render_strided(wl.cam, W, H, wl.spp, wl.bounces, 0, n) for _ in range(2)), key=lambda s: s["render_ms"])
        crc = zlib.crc32(r.read_frame(W, H).tobytes())
        r.set_tuning([4, 4, 4, 4, 4, 20, 1 | 8])
        inst = r.render_strided(wl.cam, W, H, wl.spp, wl.bounces, 0, n)
        r.set_tuning(None)
        if n == 1:
            crc0 = crc0 if crc0 is not None else crc
        print(json.dumps({"workload": wl.name, "frame": f"{W}x{H}x{wl.spp}spp", "share": f"1/{n}", "order": name,
                          "render_ms": round(plain["render_ms"], 2), "rays": plain["rays"], "tiles": plain["tiles"],
                          "tail_avg_ms": round(inst["tail_avg_ms"], 2), "tail_max_ms": round(inst["tail_max_ms"], 2),
                          "same_image_as_row_major": (crc == crc0) if n == 1 else None}), flush=True)
r.set_tile_order(None)
r.close()
