#!/usr/bin/env python
"""Scene ingest (the step before the path): host-expanded Triangle records vs indexed arrays
assembled on the device.  Prints one JSON line per workload.

    python tools/bench_ingest.py [c3 c4 c5]
"""
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import ripoff_raytracer_b200 as rr  # noqa: E402
from ripoff_raytracer_b200 import workloads  # noqa: E402


def indexed_from_triangles(t):
    """Corner-indexed arrays with shared vertices merged (what an OBJ file of the mesh holds)."""
    P = np.concatenate([t["posA"][:, :3], t["posB"][:, :3], t["posC"][:, :3]])
    N = np.concatenate([t["normalA"][:, :3], t["normalB"][:, :3], t["normalC"][:, :3]])
    up, ip = np.unique(P, axis=0, return_inverse=True)
    un, inn = np.unique(N, axis=0, return_inverse=True)
    n = len(t)
    cor = np.stack([ip[:n], ip[n:2 * n], ip[2 * n:], inn[:n], inn[n:2 * n], inn[2 * n:]], 1).astype(np.uint32)
    return up.astype(np.float32), un.astype(np.float32), cor


for name in (sys.argv[1:] or ["c3", "c4"]):
    wl = workloads.WORKLOADS[name](width=64, height=36, spp=1)
    t, m, r, sp = wl.scene.arrays()
    pos, nrm, cor = indexed_from_triangles(t)
    assert rr.triangles_from_indexed(pos, nrm, cor).tobytes() == t.tobytes()
    ren = rr.Renderer()
    res = {}
    for label, fn, nbytes in (("triangles", lambda: ren.upload_arrays(t, m, r, sp), t.nbytes),
                              ("indexed", lambda: ren.upload_indexed(pos, nrm, cor, m, r, sp), pos.nbytes + nrm.nbytes + cor.nbytes)):
        fn()  # first call allocates
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            fn()
            best = min(best, time.perf_counter() - t0)
        img = ren.render_plain(wl.cam, wl.width, wl.height, 1, 4)
        res[label] = {"upload_build_ms": round(best * 1e3, 2), "h2d_bytes": int(nbytes), "image_sum": int(img.astype(np.int64).sum())}
    ren.close()
    assert res["triangles"]["image_sum"] == res["indexed"]["image_sum"]
    print(json.dumps({"workload": wl.name, "triangles": len(t), "positions": len(pos), "normals": len(nrm), **res}), flush=True)
