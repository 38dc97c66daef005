"""The C-ABI library loads on a CPU-only box and exports every symbol include/rr_api.h declares."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from ripoff_raytracer_b200 import _abi

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (ROOT / "include" / "rr_api.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rr_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    l = _abi.lib()
    names = declared_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(l, n), f"{n} declared in rr_api.h but not exported"
        assert n in _abi.SYMBOLS, f"{n} has no ctypes prototype"


def test_wire_struct_sizes_match_reference_layout():
    # sizes/offsets of reference src/readobj.hpp:15-89 (verified on both sides)
    assert _abi.TRIANGLE.itemsize == 96
    assert _abi.MESH.itemsize == 112 and _abi.MESH.fields["material"][1] == 48 and _abi.MESH.fields["scale"][1] == 44
    assert _abi.MATERIAL.itemsize == 64 and _abi.MATERIAL.fields["emissionStrength"][1] == 48
    assert _abi.CAMERA.itemsize == 48 and _abi.CAMERA.fields["fov"][1] == 28
    assert _abi.GPU_NODE.itemsize == 48 and _abi.REF_NODE.itemsize == 64
    assert _abi.SPHERE.itemsize == 96 and _abi.SPHERE.fields["material"][1] == 32


def test_error_strings_and_version():
    l = _abi.lib()
    assert l.rr_version() >= 100
    assert l.rr_error_string(0) == b"success"
    assert b"device" in l.rr_error_string(2)


def test_no_cpu_fallback_without_device():
    """Without a CUDA device rr_create must fail loudly (reference: 'Failed to select a usable device')."""
    l = _abi.lib()
    n = C.c_int(-1)
    st = l.rr_device_count(C.byref(n))
    if st == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    assert l.rr_create(None, 0, C.byref(h)) == 2  # RR_ERR_NO_DEVICE
    assert not h.value


def test_default_camera_matches_settings():
    import ripoff_raytracer_b200 as rr

    cam = rr.default_camera(512, 512)
    assert cam["position"][0, :3].tolist() == [0.0, 150.0, 250.0]
    assert cam["yaw"][0] == np.float32(3.14) and cam["fov"][0] == 90.0 and cam["aspectRatio"][0] == 1.0


def test_new_entry_points_reject_bad_arguments_without_a_device():
    """Argument checks of the progressive / video entry points run before any CUDA call."""
    l = _abi.lib()
    cam = np.zeros(1, _abi.CAMERA)
    img = np.zeros((2, 2, 4), np.uint8)
    assert l.rr_update_meshes(None, None, 0) == 1
    assert l.rr_accum_reset(None, 4, 4) == 1
    assert l.rr_accum_add_frame(None, _abi.ptr(cam), 2, 2, 1, 1, 1, 0, _abi.ptr(img), None) == 1
    assert l.rr_accum_frame_count(None, None) == 1
    assert l.rr_accum_last_ms(None, None) == 1
    assert l.rr_render_progressive(None, _abi.ptr(cam), 2, 2, 1, 1, 1, 1, 0, None, None) == 1       # null output image
    assert l.rr_render_progressive(None, _abi.ptr(cam), 2, 2, 1, 1, 1, 0, 0, _abi.ptr(img), None) == 1  # zero frames
    assert l.rr_render_progressive(None, _abi.ptr(cam), 2, 2, 1, 1, 1, 2, 0, _abi.ptr(img), None) == 1  # null context
    assert l.rr_video_frame_setup(None, 0, 0, 1) == 1
    buf = C.create_string_buffer(8)
    assert l.rr_video_frame_path(b"a_long_directory_name", 1, buf, len(buf)) == 1                    # does not fit
    assert b"argument" in l.rr_error_string(1)
