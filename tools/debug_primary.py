import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import ripoff_raytracer_b200 as rr
from oracle.pyoracle import Oracle
g = dict(np.load(Path(__file__).resolve().parents[1] / "tests/golden/default_small.npz"))
W, H = int(g["W"]), int(g["H"])
r = rr.Renderer((0,))
r.upload_arrays(g["tris"], g["meshes"], g["ranges"])
mesh, prim, dst = r.primary_hits(g["cam"], W, H)
om, op, od = Oracle(g["tris"], g["meshes"], g["ranges"]).primary(g["cam"], W, H)
bad = (mesh != om) | (prim != op) | (dst.view(np.uint32) != od.view(np.uint32))
print("mismatch", bad.sum(), "of", bad.size, "mesh", (mesh != om).sum(), "prim", (prim != op).sum(), "dst", (dst.view(np.uint32) != od.view(np.uint32)).sum())
ys, xs = np.nonzero(bad)
for y, x in list(zip(ys, xs))[:20]:
    print(y, x, "gpu", mesh[y, x], prim[y, x], dst[y, x], "oracle", om[y, x], op[y, x], od[y, x])
print("ranges", g["ranges"])
print(g["meshes"]["material"]["type"], g["meshes"]["scale"], g["meshes"]["yaw"])
