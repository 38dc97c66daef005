"""TEST INFRASTRUCTURE -- ctypes wrappers of the checkers.

  Oracle      oracle/liboracle.so        our C restatement (oracle/rr_oracle.c)
  Reference   oracle/_ref/libref_*.so    the reference's OWN sources compiled for
                                         the host CPU (oracle/ref_shim, Makefile `ref`)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product never does.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ORACLE_SO = HERE / "liboracle.so"
REF_DIR = HERE / "_ref"

_vp, _u32, _i32, _u64 = C.c_void_p, C.c_uint32, C.c_int32, C.c_uint64


def build(ref: bool = True) -> None:
    subprocess.run(["make", "-s", "-C", str(HERE), "oracle"], check=True)
    if ref and Path("/root/reference/src/Trace.cl").exists():
        subprocess.run(["make", "-s", "-C", str(HERE), "ref"], check=True)


def _p(a):
    return C.c_void_p(0) if a is None else C.c_void_p(a.ctypes.data)


class Oracle:
    """liboracle.so: restatement of the path + our LBVH (Oracle A' / Oracle B)."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            if not ORACLE_SO.exists():
                build(ref=False)
            l = C.CDLL(str(ORACLE_SO))
            l.rro_scene_create.restype = _vp
            l.rro_scene_create.argtypes = [_vp, _u64, _vp, _vp, _u64, _vp, _u64]
            l.rro_scene_destroy.argtypes = [_vp]
            l.rro_scene_set_ref_nodes.argtypes = [_vp, _vp]
            l.rro_scene_set_brute_force.argtypes = [_vp, C.c_int]
            l.rro_lbvh_size.restype = _u64
            l.rro_lbvh_size.argtypes = [_vp, C.c_int]
            l.rro_lbvh_depth.restype = _u32
            l.rro_lbvh_depth.argtypes = [_vp, C.c_int]
            l.rro_lbvh_read.argtypes = [_vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp]
            l.rro_render.argtypes = [_vp, _vp, _u32, _u32, _u32, _u32, _i32, _vp, _vp, _vp, C.c_int]
            l.rro_primary.argtypes = [_vp, _vp, _u32, _u32, _vp, _vp, _vp, C.c_int]
            l.rro_make_seed.restype = _u32
            l.rro_make_seed.argtypes = [_u32, _i32, _u32]
            l.rro_random_value.restype = C.c_float
            l.rro_random_value.argtypes = [C.POINTER(_u32)]
            l.rro_rand01.restype = C.c_float
            l.rro_rand01.argtypes = [C.POINTER(_u32)]
            l.rro_random_direction.argtypes = [C.POINTER(_u32), _vp]
            l.rro_math.argtypes = [C.c_int, _vp, _vp, _vp, _u64]
            l.rro_make_ray.argtypes = [_vp, _u32, _u32, _u32, _u32, _vp]
            cls._lib = l
        return cls._lib

    def __init__(self, tris, meshes, ranges, spheres=None, ref_gpunodes=None):
        l = self.lib()
        self.tris = np.ascontiguousarray(tris)
        self.meshes = np.ascontiguousarray(meshes)
        self.ranges = np.ascontiguousarray(ranges)
        self.spheres = None if spheres is None or len(spheres) == 0 else np.ascontiguousarray(spheres)
        ns = 0 if self.spheres is None else len(self.spheres)
        self.h = l.rro_scene_create(_p(self.tris), len(self.tris), _p(self.meshes), _p(self.ranges), len(self.meshes),
                                    _p(self.spheres), ns)
        self._nodes = None
        if ref_gpunodes is not None:
            self._nodes = np.ascontiguousarray(ref_gpunodes)
            l.rro_scene_set_ref_nodes(self.h, _p(self._nodes))

    def close(self):
        if self.h:
            self.lib().rro_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def brute_force(self, on=True):
        """Closest hit by testing every primitive of a mesh (no hierarchy): the definition of the result."""
        self.lib().rro_scene_set_brute_force(self.h, 1 if on else 0)
        return self

    def render(self, cam, W, H, spp, bounces, frame_index=0, radiance=False, threads=8):
        rgba = np.zeros((H, W, 4), np.uint8)
        rad = np.zeros((H, W, 3), np.float32) if radiance else None
        stats = np.zeros(4, np.uint64)
        cam = np.ascontiguousarray(cam)
        self.lib().rro_render(self.h, _p(cam), W, H, spp, bounces, frame_index, _p(rgba), _p(rad), _p(stats), threads)
        st = dict(rays=int(stats[0]), box_tests=int(stats[1]), tri_tests=int(stats[2]), sphere_tests=int(stats[3]))
        return rgba, rad, st

    def primary(self, cam, W, H, threads=8):
        mesh = np.zeros((H, W), np.int32)
        prim = np.zeros((H, W), np.int32)
        dst = np.zeros((H, W), np.float32)
        cam = np.ascontiguousarray(cam)
        self.lib().rro_primary(self.h, _p(cam), W, H, _p(mesh), _p(prim), _p(dst), threads)
        return mesh, prim, dst

    def lbvh(self, which=0):
        l = self.lib()
        n = int(l.rro_lbvh_size(self.h, which))
        m = max(n, 1)
        out = dict(codes=np.zeros(m, np.uint64), order=np.zeros(m, np.uint32), left=np.zeros(m, np.int32),
                   right=np.zeros(m, np.int32), parent=np.zeros(m, np.int32), bounds=np.zeros((m, 6), np.float32))
        l.rro_lbvh_read(self.h, which, _p(out["codes"]), _p(out["order"]), _p(out["left"]), _p(out["right"]),
                        _p(out["parent"]), _p(out["bounds"]))
        out = {k: v[:n] for k, v in out.items()}
        out["depth"] = int(l.rro_lbvh_depth(self.h, which))
        return out

    @classmethod
    def math(cls, fn, x, y=None):
        x = np.ascontiguousarray(x, np.float32)
        y = x if y is None else np.ascontiguousarray(y, np.float32)
        out = np.zeros_like(x)
        cls.lib().rro_math(fn, _p(x), _p(y), _p(out), x.size)
        return out


class Reference:
    """oracle/_ref/libref_{strict,fast}.so: the reference's own kernel text + host code on the CPU."""

    _libs: dict = {}

    @classmethod
    def available(cls, variant="strict") -> bool:
        return (REF_DIR / f"libref_{variant}.so").exists()

    def __init__(self, variant="strict"):
        if variant not in self._libs:
            l = C.CDLL(str(REF_DIR / f"libref_{variant}.so"))
            l.ref_scene_default.argtypes = [C.c_char_p]
            l.ref_scene_set.argtypes = [_vp, C.c_size_t, _vp, C.c_size_t, _vp, C.c_size_t]
            l.ref_scene_from_arrays.argtypes = [_vp, C.c_size_t, _vp, _vp, C.c_size_t]
            l.ref_count.restype = C.c_size_t
            l.ref_count.argtypes = [C.c_int]
            l.ref_copy.argtypes = [C.c_int, _vp]
            l.ref_default_camera.argtypes = [_vp, C.c_int, C.c_int]
            l.ref_default_settings.argtypes = [_vp]
            l.ref_render.argtypes = [_vp, C.c_int, C.c_int, C.c_uint, C.c_uint, C.c_int, _vp, _vp, C.c_int]
            l.ref_primary.argtypes = [_vp, C.c_int, C.c_int, _vp, _vp, C.c_int]
            l.ref_make_seed.restype = C.c_uint
            l.ref_make_seed.argtypes = [C.c_uint, C.c_int, C.c_uint]
            l.ref_random_value.restype = C.c_float
            l.ref_random_value.argtypes = [C.POINTER(C.c_uint)]
            l.ref_rand01.restype = C.c_float
            l.ref_rand01.argtypes = [C.POINTER(C.c_uint)]
            l.ref_random_direction.argtypes = [C.POINTER(C.c_uint), _vp]
            l.ref_write_bmp.argtypes = [_vp, C.c_int, C.c_int, C.c_char_p]
            l.ref_video_frame_setup.argtypes = [C.c_int, C.c_int]
            self._libs[variant] = l
        self.l = self._libs[variant]

    def scene_default(self, obj_path: str):
        """Scene assembly of the reference's main() (src/main.cpp:246-272,298-304,706)."""
        self.l.ref_scene_default(str(obj_path).encode())
        return self.arrays()

    def scene_set(self, tris, meshes, gpunodes):
        self.l.ref_scene_set(_p(tris), len(tris), _p(meshes), len(meshes), _p(gpunodes), len(gpunodes))

    def scene_from_arrays(self, tris, meshes, ranges):
        """Caller arrays + the reference's own SplitBVH (SAH) hierarchy; returns the (reordered) upload arrays."""
        tris = np.ascontiguousarray(tris)
        meshes = np.ascontiguousarray(meshes)
        ranges = np.ascontiguousarray(ranges)
        self.l.ref_scene_from_arrays(_p(tris), len(tris), _p(meshes), _p(ranges), len(meshes))
        return self.arrays()

    def arrays(self):
        from ripoff_raytracer_b200._abi import GPU_NODE, MESH, TRIANGLE

        t = np.zeros(self.l.ref_count(0), TRIANGLE)
        m = np.zeros(self.l.ref_count(1), MESH)
        n = np.zeros(self.l.ref_count(2), GPU_NODE)
        self.l.ref_copy(0, _p(t))
        self.l.ref_copy(1, _p(m))
        self.l.ref_copy(2, _p(n))
        return t, m, n

    def video_frame_setup(self, frame_index, frame_count):
        """setupNextVideoFrame (src/image.hpp:385-390) on the current mesh list; returns the mesh array."""
        assert self.l.ref_video_frame_setup(frame_index, frame_count) == 0
        return self.arrays()[1]

    def default_camera(self, W, H):
        from ripoff_raytracer_b200._abi import CAMERA

        cam = np.zeros(1, CAMERA)
        self.l.ref_default_camera(_p(cam), W, H)
        return cam

    def default_settings(self):
        s = np.zeros(5, np.uint32)
        self.l.ref_default_settings(_p(s))
        return dict(width=int(s[0]), height=int(s[1]), spp=int(s[2]), bounces=int(s[3]), tile=int(s[4]))

    def render(self, cam, W, H, spp, bounces, frame_index=0, radiance=False, threads=8):
        rgba = np.zeros((H, W, 4), np.uint8)
        rad = np.zeros((H, W, 3), np.float32) if radiance else None
        cam = np.ascontiguousarray(cam)
        self.l.ref_render(_p(cam), W, H, spp, bounces, frame_index, _p(rgba), _p(rad), threads)
        return rgba, rad

    def primary(self, cam, W, H, threads=8):
        out = np.zeros((H, W, 8), np.float32)
        flags = np.zeros((H, W), np.int32)
        cam = np.ascontiguousarray(cam)
        self.l.ref_primary(_p(cam), W, H, _p(out), _p(flags), threads)
        return out, flags

    def write_bmp(self, rgba, path):
        rgba = np.ascontiguousarray(rgba)
        H, W = rgba.shape[:2]
        self.l.ref_write_bmp(_p(rgba), W, H, str(path).encode())


def mesh_ranges_from_gpunodes(meshes, gpunodes):
    """Triangle range of each mesh = union of the leaves under its root (GPUNode layout, src/image.hpp:116-125)."""
    from ripoff_raytracer_b200._abi import MESH_RANGE

    out = np.zeros(len(meshes), MESH_RANGE)
    for i, m in enumerate(meshes):
        lo, hi = None, 0
        stack = [int(m["nodeIdx"])]
        while stack:
            n = gpunodes[stack.pop()]
            if n["numTriangles"] > 0:
                a, b = int(n["index"]), int(n["index"] + n["numTriangles"])
                lo = a if lo is None else min(lo, a)
                hi = max(hi, b)
            elif n["index"] != 0:
                stack += [int(n["index"]), int(n["index"]) + 1]
        out[i] = (lo or 0, hi - (lo or 0))
    return out
