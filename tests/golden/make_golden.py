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


def progressive_video_case():
    """Progressive averaging (src/main.cpp:481, 575-582) and the video pose (src/image.hpp:385-390) on the
    default_small scene: the reference kernel renders frame k with kernel arg 7 = k (1-based), the 8-bit frames
    are summed in integers and divided by the frame count; setupNextVideoFrame is the reference's own text with
    VIDEO_FRAME_COUNT bound to a variable (oracle/ref_shim/ref_driver.cpp)."""
    ref = Reference("strict")
    g = dict(np.load(OUT / "default_small.npz"))
    W, H = int(g["W"]), int(g["H"])
    ref.scene_set(g["tris"], g["meshes"], g["gpunodes"])
    data = {}
    spp, bounces, frames = 2, 8, 5
    sums = np.zeros((H, W, 3), np.uint32)
    for k in range(1, frames + 1):
        rgba, _ = ref.render(g["cam"], W, H, spp, bounces, frame_index=k)
        sums += rgba[..., :3]
        data[f"avg_after_{k}"] = (sums // k).astype(np.uint8)
        if k == 1:
            data["frame_1"] = rgba
    data["prog_spp"], data["prog_bounces"], data["prog_frames"] = spp, bounces, frames
    count = 6
    yaws, images = [], []
    for idx in range(count):
        meshes = ref.video_frame_setup(idx, count)
        yaws.append(meshes["yaw"][-1])
        if idx in (0, 2, 5):
            rgba, _ = ref.render(g["cam"], W, H, 2, 8, frame_index=0)
            images.append(rgba)
    data["video_count"] = count
    data["video_yaw"] = np.array(yaws, np.float32)
    data["video_rgba_idx"] = np.array([0, 2, 5])
    data["video_rgba"] = np.array(images)
    data["video_yaw_count1"] = np.array([ref.video_frame_setup(i, 1)["yaw"][-1] for i in range(4)], np.float32)
    data["video_yaw_count360"] = np.array([ref.video_frame_setup(i, 360)["yaw"][-1] for i in (0, 1, 7, 180, 359)], np.float32)
    np.savez_compressed(OUT / "progressive_video.npz", **data)
    print("progressive_video", {k: getattr(v, "shape", v) for k, v in data.items()})


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "progressive_video":  # added later: leaves the older fixtures untouched
        progressive_video_case()
        sys.exit(0)
    default_scene_case("default_small", 16, 8, 64, 64, [(1, 1), (1, 8), (4, 50), (16, 50)])
    default_scene_case("default_wide", 24, 12, 96, 54, [(1, 1), (2, 50)])
    rng_case()
    progressive_video_case()
