"""Generates tests/golden/*.npz from the REFERENCE's own code.

Run in the build container (needs /root/reference and oracle/_ref built by
`make -C oracle ref`):   python tests/golden/make_golden.py

The reference has no tests, golden images or fixtures of its own,
so these vectors are produced by running its kernel text (src/Trace.cl compiled
for the host through oracle/ref_shim, strict IEEE build) on scenes assembled by
its own host code (readobj.hpp loader + SAH builder, image.hpp Cornell box).
Stored per case: the exact upload arrays (triangleList after the reference's
SplitBVH reordering, meshList, GPUNode list), the camera, and the outputs
(8-bit image, float radiance, primary-hit records).
"""
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.pyoracle import Reference, mesh_ranges_from_gpunodes  # noqa: E402
from ripoff_raytracer_b200 import scenes  # noqa: E402

OUT = Path(__file__).resolve().parent


def default_scene_case(name, nu, nv, W, H, renders):
    ref = Reference("strict")
    with tempfile.TemporaryDirectory() as td:
        obj = Path(td) / "knight.obj"
        v, n, f = scenes.uv_sphere(nu, nv)
        scenes.write_obj(obj, v, n, f)
        obj_text = obj.read_bytes()
        tris, meshes, nodes = ref.scene_default(obj)
    cam = ref.default_camera(W, H)
    data = dict(tris=tris, meshes=meshes, gpunodes=nodes, ranges=mesh_ranges_from_gpunodes(meshes, nodes), cam=cam,
                W=W, H=H, obj_text=np.frombuffer(obj_text, np.uint8))
    hit, flags = ref.primary(cam, W, H)
    data["primary_hit"] = hit
    data["primary_flags"] = flags
    for spp, bounces in renders:
        rgba, rad = ref.render(cam, W, H, spp, bounces, radiance=True)
        data[f"rgba_s{spp}_b{bounces}"] = rgba
        data[f"rad_s{spp}_b{bounces}"] = rad
    np.savez_compressed(OUT / f"{name}.npz", **data)
    print(name, {k: getattr(v, "shape", v) for k, v in data.items()})


def rng_case():
    ref = Reference("strict")
    import ctypes as C
    pix = np.array([0, 1, 511, 262143, 2073599, 33177599], np.uint32)
    seeds = np.array([ref.l.ref_make_seed(int(p), 0, 0) for p in pix], np.uint32)
    rv_state, rv_float, r01_float, dirs = [], [], [], []
    for s in seeds:
        st = C.c_uint(int(s))
        fl = [ref.l.ref_random_value(C.byref(st)) for _ in range(4)]
        rv_float.append(fl)
        rv_state.append(st.value)
        st2 = C.c_uint(int(s))
        r01_float.append([ref.l.ref_rand01(C.byref(st2)) for _ in range(4)])
        st3 = C.c_uint(int(s))
        d = np.zeros(3, np.float32)
        ref.l.ref_random_direction(C.byref(st3), C.c_void_p(d.ctypes.data))
        dirs.append(d)
    np.savez_compressed(OUT / "rng.npz", pixels=pix, seeds=seeds, rv_state_after4=np.array(rv_state, np.uint32),
                        rv_float=np.array(rv_float, np.float32), r01_float=np.array(r01_float, np.float32),
                        random_direction=np.array(dirs, np.float32))
    print("rng", seeds)


if __name__ == "__main__":
    default_scene_case("default_small", 16, 8, 64, 64, [(1, 1), (1, 8), (4, 50), (16, 50)])
    default_scene_case("default_wide", 24, 12, 96, 54, [(1, 1), (2, 50)])
    rng_case()
