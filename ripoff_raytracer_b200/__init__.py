"""ripoff_raytracer_b200 -- B200-native render hot path of ripoff-raytracer.

The product is the C-ABI library `csrc/librr_b200.so` (include/rr_api.h): a
GPU-built LBVH plus hand-written sm_100a path-tracing kernels behind the
reference's host-driver boundary (reference src/image.hpp).  This package is
the thin Python mirror of that boundary used by the tests and bench.py.
"""
from . import _abi, scenes  # noqa: F401
from .api import (Renderer, Scene, default_camera, default_scene, load_obj_indexed, triangles_from_indexed,  # noqa: F401
                  video_frame_path, video_frame_setup, write_bmp)

__all__ = ["Renderer", "Scene", "default_camera", "default_scene", "load_obj_indexed", "triangles_from_indexed",
           "video_frame_path", "video_frame_setup", "write_bmp", "scenes"]
