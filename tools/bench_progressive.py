#!/usr/bin/env python
"""progressive accumulation and video re-posing.  One JSON line per measurement.

  accumulate   k_accum_add (integer sums + running average, src/main.cpp:575-582 on the device): device time and
               achieved HBM bandwidth (32 algorithmic bytes per pixel) against MEASURED_PEAKS.json
  video        per-frame scene change: rr_update_meshes (re-pose, LBVH kept) against a full rr_upload_scene
               (what the reference's loop does: generateBuffers per frame, src/main.cpp:691-693)

    python tools/bench_progressive.py [--workload c4]
"""
import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import ripoff_raytracer_b200 as rr  # noqa: E402
from ripoff_raytracer_b200 import workloads  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c4")
a = ap.parse_args()
peak = 6650.0
try:
    peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]
except Exception:
    pass

wl = workloads.WORKLOADS[a.workload](width=64, height=36, spp=1)
r = rr.Renderer()
r.upload(wl.scene)
for W, H in ((1920, 1080), (3840, 2160), (7680, 4320)):
    cam = wl.cam.copy()
    cam["aspectRatio"] = W / H
    r.accum_reset(W, H)
    ms = []
    for k in range(1, 8):
        r.accum_add_frame(cam, W, H, 1, 1, frame_index=k, want_average=False)  # 1 spp, 1 bounce: the render is short
        ms.append(r.accum_last_ms())
    best = min(ms[2:])
    gbs = W * H * 32 / best / 1e6
    print(json.dumps({"kernel": "k_accum_add", "frame": f"{W}x{H}", "ms": round(best, 4), "bytes": W * H * 32, "GB_s": round(gbs, 1),
                      "hbm_peak_GB_s": peak, "frac": round(gbs / peak, 3)}), flush=True)

t, m, rg, sp = wl.scene.arrays()
best_up = best_pose = 1e9
for k in range(4):
    rr.video_frame_setup(m, k, 60)
    t0 = time.perf_counter()
    r.upload_arrays(t, m, rg, sp)
    best_up = min(best_up, time.perf_counter() - t0)
    t0 = time.perf_counter()
    r.update_meshes(m)
    best_pose = min(best_pose, time.perf_counter() - t0)
print(json.dumps({"workload": wl.name, "triangles": int(len(t)), "meshes": int(len(m)), "rr_upload_scene_ms": round(best_up * 1e3, 3),
                  "rr_update_meshes_ms": round(best_pose * 1e3, 3), "h2d_bytes_upload": int(t.nbytes + m.nbytes + sp.nbytes),
                  "h2d_bytes_update": int(m.nbytes)}), flush=True)
r.close()
