"""ctypes binding of include/rr_api.h (librr_b200.so).

The wire structs are numpy structured dtypes whose offsets equal the
reference's host structs (reference src/readobj.hpp:15-89; sizes verified on
both sides in tests/test_abi.py).  There is no CPU fallback: if the CUDA
library is missing or cannot be loaded, `lib()` raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "csrc" / "librr_b200.so"

F3 = ("<f4", (4,))

MATERIAL = np.dtype(
    {
        "names": ["type", "ior", "color", "emissionColor", "emissionStrength", "reflectiveness", "specularProbability"],
        "formats": ["<i4", "<f4", F3, F3, "<f4", "<f4", "<f4"],
        "offsets": [0, 4, 16, 32, 48, 52, 56],
        "itemsize": 64,
    }
)
TRIANGLE = np.dtype(
    {
        "names": ["posA", "posB", "posC", "normalA", "normalB", "normalC"],
        "formats": [F3] * 6,
        "offsets": [0, 16, 32, 48, 64, 80],
        "itemsize": 96,
    }
)
MESH = np.dtype(
    {
        "names": ["nodeIdx", "pos", "pitch", "yaw", "roll", "scale", "material"],
        "formats": ["<u8", F3, "<f4", "<f4", "<f4", "<f4", MATERIAL],
        "offsets": [0, 16, 32, 36, 40, 44, 48],
        "itemsize": 112,
    }
)
REF_NODE = np.dtype(
    {
        "names": ["bmin", "bmax", "childIndex", "firstTriangleIdx", "numTriangles"],
        "formats": [F3, F3, "<u8", "<u8", "<u8"],
        "offsets": [0, 16, 32, 40, 48],
        "itemsize": 64,
    }
)
GPU_NODE = np.dtype(
    {
        "names": ["bmin", "bmax", "index", "numTriangles"],
        "formats": [F3, F3, "<u8", "<u8"],
        "offsets": [0, 16, 32, 40],
        "itemsize": 48,
    }
)
CAMERA = np.dtype(
    {
        "names": ["position", "pitch", "yaw", "roll", "fov", "aspectRatio"],
        "formats": [F3, "<f4", "<f4", "<f4", "<f4", "<f4"],
        "offsets": [0, 16, 20, 24, 28, 32],
        "itemsize": 48,
    }
)
SPHERE = np.dtype(
    {
        "names": ["center", "radius", "material"],
        "formats": [F3, "<f4", MATERIAL],
        "offsets": [0, 16, 32],
        "itemsize": 96,
    }
)
MESH_RANGE = np.dtype([("firstTriangle", "<u8"), ("numTriangles", "<u8")])

MATERIAL_SOLID, MATERIAL_CHECKER, MATERIAL_INVISIBLE, MATERIAL_GLASSY, MATERIAL_ONESIDED = range(5)


class Stats(C.Structure):
    _fields_ = [
        ("samples", C.c_uint64),
        ("rays", C.c_uint64),
        ("stack_overflows", C.c_uint64),
        ("box_tests", C.c_uint64),
        ("tri_tests", C.c_uint64),
        ("sphere_tests", C.c_uint64),
        ("tiles", C.c_uint64),
        ("render_ms", C.c_float),
        ("build_ms", C.c_float),
        ("phase_runs", C.c_uint64 * 5),
        ("phase_lanes", C.c_uint64 * 5),
        ("tail_avg_ms", C.c_float),
        ("tail_max_ms", C.c_float),
    ]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_}
        d["phase_runs"] = list(d["phase_runs"])
        d["phase_lanes"] = list(d["phase_lanes"])
        return d


# every symbol include/rr_api.h declares: name -> (restype, argtypes)
_vp, _sz, _u32, _i32, _u64 = C.c_void_p, C.c_size_t, C.c_uint32, C.c_int32, C.c_uint64
SYMBOLS = {
    "rr_error_string": (C.c_char_p, [C.c_int]),
    "rr_last_error": (C.c_char_p, []),
    "rr_version": (C.c_int, []),
    "rr_create": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(_vp)]),
    "rr_destroy": (None, [_vp]),
    "rr_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "rr_device_info": (C.c_int, [C.c_int, C.c_char_p, _sz, C.POINTER(C.c_int), C.POINTER(_u64)]),
    "rr_upload_scene": (C.c_int, [_vp, _vp, _sz, _vp, _vp, _sz, _vp, _sz]),
    "rr_upload_scene_indexed": (C.c_int, [_vp, _vp, _sz, _vp, _sz, _vp, _sz, _vp, _vp, _sz, _vp, _sz]),
    "rr_obj_load": (C.c_int, [C.c_char_p, C.POINTER(_vp)]),
    "rr_obj_destroy": (None, [_vp]),
    "rr_obj_position_count": (_sz, [_vp]),
    "rr_obj_normal_count": (_sz, [_vp]),
    "rr_obj_triangle_count": (_sz, [_vp]),
    "rr_obj_positions": (_vp, [_vp]),
    "rr_obj_normals": (_vp, [_vp]),
    "rr_obj_corners": (_vp, [_vp]),
    "rr_upload_scene_ref": (C.c_int, [_vp, _vp, _sz, _vp, _sz, _vp, _sz]),
    "rr_update_meshes": (C.c_int, [_vp, _vp, _sz]),
    "rr_render": (C.c_int, [_vp, _vp, _u32, _u32, _u32, _u32, _i32, _u32, _vp]),
    "rr_render_ex": (C.c_int, [_vp, _vp, _u32, _u32, _u32, _u32, _i32, _u32, _vp, _vp, C.POINTER(Stats), C.c_int]),
    "rr_set_tuning": (C.c_int, [_vp, _vp, _sz]),
    "rr_render_device": (C.c_int, [_vp, _vp, _u32, _u32, _u32, _u32, _i32, _u32, C.POINTER(Stats)]),
    "rr_read_frame": (C.c_int, [_vp, _vp, _sz]),
    "rr_render_progress": (C.c_int, [_vp, C.POINTER(_u64), C.POINTER(_u64)]),
    "rr_probe_math": (C.c_int, [C.c_int, _vp, _vp, _vp, _u64]),
    "rr_probe_rng": (C.c_int, [_u32, _i32, _vp, _vp]),
    "rr_probe_peak": (C.c_int, [C.c_int, C.POINTER(C.c_float)]),
    "rr_accum_reset": (C.c_int, [_vp, _u32, _u32]),
    "rr_accum_add_frame": (C.c_int, [_vp, _vp, _u32, _u32, _u32, _u32, _i32, _u32, _vp, C.POINTER(Stats)]),
    "rr_accum_frame_count": (C.c_int, [_vp, C.POINTER(_u32)]),
    "rr_accum_last_ms": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "rr_render_progressive": (C.c_int, [_vp, _vp, _u32, _u32, _u32, _u32, _i32, _u32, _u32, _vp, C.POINTER(Stats)]),
    "rr_primary_hits": (C.c_int, [_vp, _vp, _u32, _u32, _vp, _vp, _vp]),
    "rr_render_cost": (C.c_int, [_vp, _vp, _u32, _u32, _u32, _u32, _vp]),
    "rr_set_tile_order": (C.c_int, [_vp, _vp, _u32]),
    "rr_bvh_size": (C.c_int, [_vp, C.c_int, C.POINTER(_u64)]),
    "rr_bvh_read": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rr_queue_export": (C.c_int, [_vp, _u32, _u32, _vp, _vp]),
    "rr_queue_import": (C.c_int, [_vp, _u32, _u32, _vp, _vp]),
    "rr_queue_reset": (C.c_int, [_vp]),
    "rr_render_shared": (C.c_int, [_vp, _vp, _u32, _u32, _u32, _u32, _i32, _u32, C.POINTER(Stats)]),
    "rr_render_strided": (C.c_int, [_vp, _vp, _u32, _u32, _u32, _u32, _i32, _u32, _u32, _u32, C.POINTER(Stats)]),
    "rr_frame_device_ptr": (C.c_int, [_vp, C.POINTER(_u64), C.POINTER(_u64)]),
    "rr_write_bmp": (C.c_int, [C.c_char_p, _vp, _u32, _u32]),
    "rr_scene_create": (C.c_int, [C.POINTER(_vp)]),
    "rr_scene_destroy": (None, [_vp]),
    "rr_scene_load_obj": (C.c_int, [_vp, C.c_char_p, _vp, _vp]),
    "rr_scene_range_bounds": (C.c_int, [_vp, _vp, _vp, _vp]),
    "rr_scene_add_quad": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rr_scene_add_cornell": (C.c_int, [_vp, _vp, _vp]),
    "rr_scene_add_mesh": (C.c_int, [_vp, _vp, _vp]),
    "rr_scene_add_triangles": (C.c_int, [_vp, _vp, _sz, _vp]),
    "rr_scene_add_sphere": (C.c_int, [_vp, _vp]),
    "rr_scene_mesh": (_vp, [_vp, _sz]),
    "rr_scene_mesh_count": (_sz, [_vp]),
    "rr_scene_triangle_count": (_sz, [_vp]),
    "rr_scene_sphere_count": (_sz, [_vp]),
    "rr_scene_triangles": (_vp, [_vp]),
    "rr_scene_meshes": (_vp, [_vp]),
    "rr_scene_ranges": (_vp, [_vp]),
    "rr_scene_spheres": (_vp, [_vp]),
    "rr_scene_upload": (C.c_int, [_vp, _vp]),
    "rr_default_camera": (None, [_vp, _u32, _u32]),
    "rr_video_frame_setup": (C.c_int, [_vp, _sz, _i32, _i32]),
    "rr_video_frame_path": (C.c_int, [C.c_char_p, _i32, C.c_char_p, _sz]),
}

_lib = None


class RRError(RuntimeError):
    def __init__(self, status: int, where: str):
        l = lib()
        msg = l.rr_error_string(status).decode()
        detail = l.rr_last_error().decode()
        super().__init__(f"{where}: {msg}" + (f" ({detail})" if detail else ""))
        self.status = status


def lib() -> C.CDLL:
    """Load librr_b200.so (built in-tree by __graft_entry__.build())."""
    global _lib
    if _lib is None:
        path = Path(os.environ.get("RR_B200_LIB", LIB_PATH))
        if not path.exists():
            raise ImportError(
                f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the render path)"
            )
        l = C.CDLL(str(path))
        for name, (res, args) in SYMBOLS.items():
            if "RR_B200_LIB" in os.environ and not hasattr(l, name):
                continue  # an older build loaded for an A/B run (tools/ab.py)
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(status: int, where: str) -> None:
    if status != 0:
        raise RRError(status, where)


def ptr(a) -> C.c_void_p:
    """void* of a numpy array (None -> NULL)."""
    if a is None:
        return C.c_void_p(0)
    return C.c_void_p(a.ctypes.data)
