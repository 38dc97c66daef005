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


def zoo_scene_arrays():
    """Every material the compiled reference terminates on (Solid diffuse / specular, Checker, Glassy, OneSided),
    meshes with pitch, yaw AND roll != 0 and scale in {0.5, 0.8, 1.2}, each mesh with its OWN triangle range (the
    reference's SplitBVH re-partitions a range in place, so two meshes cannot share one).  No Invisible mesh: the
    reference's `continue` at src/Trace.cl:502-506 neither counts a bounce nor moves the origin past the surface in
    float arithmetic, so its kernel does not return on such a scene."""
    import ripoff_raytracer_b200 as rr
    from ripoff_raytracer_b200 import _abi

    s = rr.Scene()

    def mesh(pos, pyr, scale, mtype, color, emis=(0, 0, 0), strength=0.0, refl=0.0, spec=0.0, ior=1.0):
        m = np.zeros(1, _abi.MESH)
        m["pos"][0, :3] = pos
        m["pitch"], m["yaw"], m["roll"] = pyr
        m["scale"] = scale
        mm = m["material"]
        mm["type"] = mtype
        mm["ior"] = ior
        mm["color"][0, :3] = color
        mm["emissionColor"][0, :3] = emis
        mm["emissionStrength"] = strength
        mm["reflectiveness"] = refl
        mm["specularProbability"] = spec
        return m

    # room: Checker floor (emissionStrength is the cell size, emissionColor the second colour: src/Trace.cl:509-524),
    # ceiling, three walls, a OneSided front wall the camera looks through from behind, a ceiling light
    s.add_quad((-300, 0, -300), (300, 0, -300), (300, 0, 300), (-300, 0, 300), (0, 1, 0), (0.8, 0.8, 0.8))
    fl = s.mesh(s.n_meshes - 1)["material"]
    fl["type"] = _abi.MATERIAL_CHECKER
    fl["emissionColor"][0, :3] = (0.1, 0.1, 0.3)
    fl["emissionStrength"] = 40.0
    fl["specularProbability"] = 0.3
    fl["reflectiveness"] = 0.6
    s.add_quad((-300, 300, -300), (300, 300, -300), (300, 300, 300), (-300, 300, 300), (0, -1, 0), (0.9, 0.9, 0.9))
    s.add_quad((-300, 0, -300), (300, 0, -300), (300, 300, -300), (-300, 300, -300), (0, 0, 1), (0.2, 0.7, 0.2))
    s.add_quad((-300, 0, 300), (300, 0, 300), (300, 300, 300), (-300, 300, 300), (0, 0, -1), (1.0, 1.0, 1.0))
    s.mesh(s.n_meshes - 1)["material"]["type"] = _abi.MATERIAL_ONESIDED
    s.add_quad((-300, 0, -300), (-300, 0, 300), (-300, 300, 300), (-300, 300, -300), (1, 0, 0), (0.2, 0.2, 0.9))
    s.add_quad((300, 0, -300), (300, 0, 300), (300, 300, 300), (300, 300, -300), (-1, 0, 0), (0.9, 0.2, 0.2))
    s.add_quad((-80, 299, -80), (80, 299, -80), (80, 299, 80), (-80, 299, 80), (0, -1, 0), (1, 1, 1))
    lm = s.mesh(s.n_meshes - 1)["material"]
    lm["emissionColor"][0, :3] = 1.0
    lm["emissionStrength"] = 6.0
    poses = [
        ((-120, 70, -40), (0.3, 1.1, -0.4), 1.2, _abi.MATERIAL_SOLID, (0.9, 0.6, 0.3), dict(refl=0.3, spec=0.5)),
        ((110, 80, 20), (-0.7, 2.5, 0.9), 0.8, _abi.MATERIAL_GLASSY, (0.9, 0.9, 0.9), dict(ior=1.5)),
        ((0, 60, -120), (0.5, 0.4, 2.0), 1.2, _abi.MATERIAL_SOLID, (0.8, 0.8, 0.8), dict(refl=1.0, spec=1.0)),
        ((-40, 200, 0), (0.2, -1.2, 0.7), 0.5, _abi.MATERIAL_ONESIDED, (0.5, 0.9, 0.9), {}),
        ((60, 150, -60), (-1.0, 0.3, -2.2), 0.5, _abi.MATERIAL_GLASSY, (0.7, 0.9, 0.7), dict(ior=1.33)),
        ((-150, 40, 120), (2.8, -0.6, 0.1), 0.8, _abi.MATERIAL_CHECKER, (0.9, 0.9, 0.2), dict(emis=(0.2, 0.1, 0.6), strength=15.0)),
    ]
    for k, (pos, pyr, scale, mtype, color, kw) in enumerate(poses):
        if k % 2 == 0:
            v, n, f = scenes.displaced_icosphere(2, radius=60.0, center=(0.0, 0.0, 0.0), seed=20 + k)
        else:
            v, n, f = scenes.uv_sphere(14, 7, radius=45.0, center=(0.0, 0.0, 0.0))
        rng = s.add_triangles(scenes.mesh_triangles(v, n, f))  # a fresh range for every mesh
        s.add_mesh(mesh(pos, pyr, scale, mtype, color, **kw), rng)
    t, m, r, _ = s.arrays()
    from ripoff_raytracer_b200 import _abi as A
    cam = np.zeros(1, A.CAMERA)
    return t, m, r, cam


def zoo_case(W=96, H=72, renders=((1, 1), (1, 12), (4, 50))):
    """Reference-generated fixture for the material / rotation code no default-scene fixture reaches
    (src/Trace.cl:509-558 Checker + Glassy, :219-236 refract / reflect, :401-432 CalculateReflectance, :90-100
    makeRotation with all three angles)."""
    ref = Reference("strict")
    t, m, r, cam = zoo_scene_arrays()
    cam["position"][0, :3] = (30.0, 140.0, 280.0)
    cam["pitch"], cam["yaw"], cam["roll"] = 0.25, 3.0, 0.05
    cam["fov"] = 80.0
    cam["aspectRatio"] = np.float32(W) / np.float32(H)
    tris, meshes, nodes = ref.scene_from_arrays(t, m, r)  # triangles come back re-ordered by the reference's SplitBVH
    data = dict(tris=tris, meshes=meshes, gpunodes=nodes, ranges=mesh_ranges_from_gpunodes(meshes, nodes), cam=cam, W=W, H=H)
    assert np.array_equal(data["ranges"]["firstTriangle"], r["firstTriangle"]) and np.array_equal(data["ranges"]["numTriangles"], r["numTriangles"])
    hit, flags = ref.primary(cam, W, H)
    data["primary_hit"] = hit
    data["primary_flags"] = flags
    for spp, bounces in renders:
        rgba, rad = ref.render(cam, W, H, spp, bounces, radiance=True)
        data[f"rgba_s{spp}_b{bounces}"] = rgba
        data[f"rad_s{spp}_b{bounces}"] = rad
    np.savez_compressed(OUT / "zoo_ref.npz", **data)
    types = sorted(set((flags[flags >= 0] >> 8).tolist()))
    print("zoo_ref", {k: getattr(v, "shape", v) for k, v in data.items()}, "material types seen by primary rays:", types)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "progressive_video":  # added later: leaves the older fixtures untouched
        progressive_video_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "zoo":
        zoo_case()
        sys.exit(0)
    default_scene_case("default_small", 16, 8, 64, 64, [(1, 1), (1, 8), (4, 50), (16, 50)])
    default_scene_case("default_wide", 24, 12, 96, 54, [(1, 1), (2, 50)])
    rng_case()
    progressive_video_case()
    zoo_case()
