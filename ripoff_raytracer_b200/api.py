"""Host-side mirror of the reference's driver interface over the C ABI.

Names follow the reference: a *scene* is the triangleList / meshList pair of
src/readobj.hpp:91-94, `Scene.load_obj` is loadMeshFromOBJFile, `add_quad` is
addQuad, `add_cornell` is addCornellBoxToScene (src/image.hpp:401-448);
`Renderer` stands where generateKernelForDevice / generateBuffers /
singleThreadedCompute / multiThreadedCompute stand (src/image.hpp:30-381).
Every call goes to librr_b200.so; nothing here computes pixels.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from ._abi import CAMERA, MESH, MESH_RANGE, SPHERE, TRIANGLE, Stats, check, lib, ptr


def _f3(v):
    return np.ascontiguousarray(np.asarray(v, dtype=np.float32).reshape(3))


def default_camera(width: int, height: int) -> np.ndarray:
    """CameraInformation of src/main.cpp:299-304 (start pose of src/settings.hpp:23-28, fov 90)."""
    cam = np.zeros(1, CAMERA)
    lib().rr_default_camera(ptr(cam), width, height)
    return cam


def video_frame_setup(meshes: np.ndarray, frame_index: int, frame_count: int) -> None:
    """setupNextVideoFrame (src/image.hpp:385-390): edits the LAST mesh's yaw in place."""
    assert meshes.dtype == MESH and meshes.flags["C_CONTIGUOUS"]
    check(lib().rr_video_frame_setup(ptr(meshes), len(meshes), frame_index, frame_count), "rr_video_frame_setup")


def video_frame_path(directory, frame_number: int) -> str:
    """<dir>/output_<n>.bmp (src/main.cpp:701; the pattern render.sh hands to ffmpeg)."""
    buf = C.create_string_buffer(4096)
    check(lib().rr_video_frame_path(str(directory).encode(), frame_number, buf, len(buf)), "rr_video_frame_path")
    return buf.value.decode()


def write_bmp(path, rgba: np.ndarray) -> None:
    """placeImageDataIntoBMP (src/math.hpp:117-164)."""
    rgba = np.ascontiguousarray(rgba, np.uint8)
    h, w = rgba.shape[:2]
    check(lib().rr_write_bmp(str(path).encode(), ptr(rgba), w, h), "rr_write_bmp")


class Scene:
    """Scene builder (triangleList + meshList + spheres) held by the C++ host layer."""

    def __init__(self):
        h = C.c_void_p()
        check(lib().rr_scene_create(C.byref(h)), "rr_scene_create")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            lib().rr_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_obj(self, path):
        mesh = np.zeros(1, MESH)
        rng = np.zeros(1, MESH_RANGE)
        check(lib().rr_scene_load_obj(self.h, str(path).encode(), ptr(mesh), ptr(rng)), f"rr_scene_load_obj({path})")
        return mesh, rng

    def add_triangles(self, tris: np.ndarray):
        tris = np.ascontiguousarray(tris, TRIANGLE)
        rng = np.zeros(1, MESH_RANGE)
        check(lib().rr_scene_add_triangles(self.h, ptr(tris), len(tris), ptr(rng)), "rr_scene_add_triangles")
        return rng

    def add_mesh(self, mesh: np.ndarray, rng: np.ndarray):
        mesh = np.ascontiguousarray(mesh, MESH)
        rng = np.ascontiguousarray(rng, MESH_RANGE)
        check(lib().rr_scene_add_mesh(self.h, ptr(mesh), ptr(rng)), "rr_scene_add_mesh")

    def add_quad(self, a, b, c, d, normal, color):
        args = [_f3(v) for v in (a, b, c, d, normal, color)]
        check(lib().rr_scene_add_quad(self.h, *[ptr(v) for v in args]), "rr_scene_add_quad")

    def add_cornell(self, mesh: np.ndarray, rng: np.ndarray):
        mesh = np.ascontiguousarray(mesh, MESH)
        rng = np.ascontiguousarray(rng, MESH_RANGE)
        check(lib().rr_scene_add_cornell(self.h, ptr(mesh), ptr(rng)), "rr_scene_add_cornell")

    def add_spheres(self, spheres: np.ndarray):
        spheres = np.ascontiguousarray(spheres, SPHERE)
        for i in range(len(spheres)):
            check(lib().rr_scene_add_sphere(self.h, ptr(spheres[i:i + 1])), "rr_scene_add_sphere")

    def range_bounds(self, rng: np.ndarray):
        rng = np.ascontiguousarray(rng, MESH_RANGE)
        lo, hi = np.zeros(3, np.float32), np.zeros(3, np.float32)
        check(lib().rr_scene_range_bounds(self.h, ptr(rng), ptr(lo), ptr(hi)), "rr_scene_range_bounds")
        return lo, hi

    def mesh(self, index: int) -> np.ndarray:
        """Writable view of mesh `index` (the reference edits meshList.back() in place)."""
        p = lib().rr_scene_mesh(self.h, index)
        if not p:
            raise IndexError(index)
        buf = (C.c_char * MESH.itemsize).from_address(p)
        return np.frombuffer(buf, dtype=MESH, count=1)

    @property
    def n_meshes(self):
        return lib().rr_scene_mesh_count(self.h)

    @property
    def n_triangles(self):
        return lib().rr_scene_triangle_count(self.h)

    @property
    def n_spheres(self):
        return lib().rr_scene_sphere_count(self.h)

    def arrays(self):
        """Copies of (triangles, meshes, ranges, spheres)."""
        l = lib()

        def copy(p, n, dt):
            if n == 0:
                return np.zeros(0, dt)
            buf = (C.c_char * (n * dt.itemsize)).from_address(p)
            return np.frombuffer(buf, dtype=dt, count=n).copy()

        return (copy(l.rr_scene_triangles(self.h), self.n_triangles, TRIANGLE),
                copy(l.rr_scene_meshes(self.h), self.n_meshes, MESH),
                copy(l.rr_scene_ranges(self.h), self.n_meshes, MESH_RANGE),
                copy(l.rr_scene_spheres(self.h), self.n_spheres, SPHERE))


def load_obj_indexed(path):
    """rr_obj_load: OBJ text -> (positions (n,3) f32, normals (m,3) f32, corners (t,6) u32: v0 v1 v2 n0 n1 n2)."""
    h = C.c_void_p()
    check(lib().rr_obj_load(str(path).encode(), C.byref(h)), "rr_obj_load")
    try:
        l = lib()
        npos, nnrm, ntri = l.rr_obj_position_count(h), l.rr_obj_normal_count(h), l.rr_obj_triangle_count(h)

        def copy(p, n, dt):
            if n == 0:
                return np.zeros(0, dt)
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), (n * np.dtype(dt).itemsize,)).view(dt).copy()

        pos = copy(l.rr_obj_positions(h), npos * 3, np.float32).reshape(-1, 3)
        nrm = copy(l.rr_obj_normals(h), nnrm * 3, np.float32).reshape(-1, 3)
        cor = copy(l.rr_obj_corners(h), ntri * 6, np.uint32).reshape(-1, 6)
        return pos, nrm, cor
    finally:
        lib().rr_obj_destroy(h)


def triangles_from_indexed(positions, normals, corners) -> np.ndarray:
    """Host-side expansion of indexed arrays to Triangle records (what the device gather produces)."""
    t = np.zeros(len(corners), TRIANGLE)
    for k, name in enumerate(("posA", "posB", "posC")):
        t[name][:, :3] = positions[corners[:, k]]
    for k, name in enumerate(("normalA", "normalB", "normalC")):
        t[name][:, :3] = normals[corners[:, 3 + k]]
    return t


def default_scene(obj_path) -> Scene:
    """Scene assembly of the reference's main() (src/main.cpp:246-272, 298, 706):
    OBJ mesh (Solid white, specularProbability 1, scale 0.5), Cornell box around it,
    OBJ mesh appended LAST, then yaw = 5.5 (setupNextVideoFrame, src/image.hpp:385-390)."""
    s = Scene()
    mesh, rng = s.load_obj(obj_path)
    m = mesh["material"]
    m["type"] = _abi.MATERIAL_SOLID
    m["ior"] = 1.0
    m["color"][:, :3] = 1.0
    m["emissionColor"][:] = 0.0
    m["emissionStrength"] = 0.0
    m["reflectiveness"] = 0.0
    m["specularProbability"] = 1.0
    mesh["scale"] = 0.5
    s.add_cornell(mesh, rng)
    s.add_mesh(mesh, rng)
    s.mesh(s.n_meshes - 1)["yaw"] = np.float32(0.0) + np.float32(5.5)
    return s


class Renderer:
    """One render context over one or more CUDA devices of this process."""

    def __init__(self, devices=(0,)):
        ords = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        check(lib().rr_create(ords, len(devices), C.byref(h)), "rr_create")
        self.h = h
        self._keep = None

    def close(self):
        if getattr(self, "h", None):
            lib().rr_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- generateBuffers ----------------------------------------------------
    def upload(self, scene: Scene):
        check(lib().rr_scene_upload(self.h, scene.h), "rr_scene_upload")

    def upload_arrays(self, tris, meshes, ranges, spheres=None):
        tris = np.ascontiguousarray(tris, TRIANGLE)
        meshes = np.ascontiguousarray(meshes, MESH)
        ranges = np.ascontiguousarray(ranges, MESH_RANGE)
        ns = 0 if spheres is None else len(spheres)
        spheres = None if ns == 0 else np.ascontiguousarray(spheres, SPHERE)
        check(lib().rr_upload_scene(self.h, ptr(tris), len(tris), ptr(meshes), ptr(ranges), len(meshes), ptr(spheres), ns),
              "rr_upload_scene")

    def upload_indexed(self, positions, normals, corners, meshes, ranges, spheres=None):
        """rr_upload_scene_indexed: raw OBJ arrays in, the Triangle records are assembled on the device."""
        positions = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
        normals = np.ascontiguousarray(normals, np.float32).reshape(-1, 3)
        corners = np.ascontiguousarray(corners, np.uint32).reshape(-1, 6)
        meshes = np.ascontiguousarray(meshes, MESH)
        ranges = np.ascontiguousarray(ranges, MESH_RANGE)
        spheres = np.zeros(0, SPHERE) if spheres is None else np.ascontiguousarray(spheres, SPHERE)
        check(lib().rr_upload_scene_indexed(self.h, ptr(positions), len(positions), ptr(normals), len(normals), ptr(corners),
                                            len(corners), ptr(meshes), ptr(ranges), len(meshes),
                                            ptr(spheres) if len(spheres) else None, len(spheres)), "rr_upload_scene_indexed")

    def upload_ref(self, tris, meshes, ref_nodes):
        """generateBuffers' own argument list: triangleList, meshList, nodeList (host Node layout)."""
        tris = np.ascontiguousarray(tris, TRIANGLE)
        meshes = np.ascontiguousarray(meshes, MESH)
        ref_nodes = np.ascontiguousarray(ref_nodes, _abi.REF_NODE)
        check(lib().rr_upload_scene_ref(self.h, ptr(tris), len(tris), ptr(meshes), len(meshes), ptr(ref_nodes), len(ref_nodes)),
              "rr_upload_scene_ref")

    def update_meshes(self, meshes):
        """rr_update_meshes: new MeshInfo (pose / material) for the uploaded scene, no re-upload, no rebuild."""
        meshes = np.ascontiguousarray(meshes, MESH)
        check(lib().rr_update_meshes(self.h, ptr(meshes), len(meshes)), "rr_update_meshes")

    # -- progressive mode (src/main.cpp:481, 575-582) ---------------------------------
    def accum_reset(self, width, height):
        check(lib().rr_accum_reset(self.h, width, height), "rr_accum_reset")

    def accum_add_frame(self, cam, width, height, spp, bounces, frame_index, tile=0, want_average=True):
        cam = np.ascontiguousarray(cam, CAMERA)
        avg = np.zeros((height, width, 4), np.uint8) if want_average else None
        st = Stats()
        check(lib().rr_accum_add_frame(self.h, ptr(cam), width, height, spp, bounces, frame_index, tile, ptr(avg), C.byref(st)),
              "rr_accum_add_frame")
        return avg, st.as_dict()

    def accum_last_ms(self) -> float:
        ms = C.c_float()
        check(lib().rr_accum_last_ms(self.h, C.byref(ms)), "rr_accum_last_ms")
        return float(ms.value)

    def accum_frame_count(self) -> int:
        n = C.c_uint32()
        check(lib().rr_accum_frame_count(self.h, C.byref(n)), "rr_accum_frame_count")
        return int(n.value)

    def render_progressive(self, cam, width, height, spp, bounces, n_frames, first_frame_index=1, tile=0):
        cam = np.ascontiguousarray(cam, CAMERA)
        rgba = np.zeros((height, width, 4), np.uint8)
        st = Stats()
        check(lib().rr_render_progressive(self.h, ptr(cam), width, height, spp, bounces, first_frame_index, n_frames, tile,
                                          ptr(rgba), C.byref(st)), "rr_render_progressive")
        return rgba, st.as_dict()

    # -- singleThreadedCompute / multiThreadedCompute -----------------------------
    def render(self, cam, width, height, spp, bounces, frame_index=0, tile=0, radiance=False, count_tests=False,
               out=None):
        cam = np.ascontiguousarray(cam, CAMERA)
        rgba = out if out is not None else np.zeros((height, width, 4), np.uint8)
        rad = np.zeros((height, width, 3), np.float32) if radiance else None
        st = Stats()
        check(lib().rr_render_ex(self.h, ptr(cam), width, height, spp, bounces, frame_index, tile, ptr(rgba), ptr(rad),
                                 C.byref(st), 1 if count_tests else 0), "rr_render_ex")
        return rgba, rad, st.as_dict()

    def set_tuning(self, values=None):
        """Scheduler knobs of the render kernel (rr_set_tuning); None restores the defaults."""
        if values is None:
            check(lib().rr_set_tuning(self.h, None, 0), "rr_set_tuning")
        else:
            v = np.ascontiguousarray(values, np.uint32)
            check(lib().rr_set_tuning(self.h, ptr(v), len(v)), "rr_set_tuning")

    def render_plain(self, cam, width, height, spp, bounces, frame_index=0, tile=0, out=None):
        """The boundary call itself: rr_render (host buffers in and out)."""
        cam = np.ascontiguousarray(cam, CAMERA)
        rgba = out if out is not None else np.zeros((height, width, 4), np.uint8)
        check(lib().rr_render(self.h, ptr(cam), width, height, spp, bounces, frame_index, tile, ptr(rgba)), "rr_render")
        return rgba

    def render_device(self, cam, width, height, spp, bounces, frame_index=0, tile=0):
        cam = np.ascontiguousarray(cam, CAMERA)
        st = Stats()
        check(lib().rr_render_device(self.h, ptr(cam), width, height, spp, bounces, frame_index, tile, C.byref(st)),
              "rr_render_device")
        return st.as_dict()

    def read_frame(self, width, height, out=None):
        rgba = out if out is not None else np.empty((height, width, 4), np.uint8)
        check(lib().rr_read_frame(self.h, ptr(rgba), rgba.nbytes), "rr_read_frame")
        return rgba

    def primary_hits(self, cam, width, height):
        cam = np.ascontiguousarray(cam, CAMERA)
        mesh = np.zeros((height, width), np.int32)
        prim = np.zeros((height, width), np.int32)
        dst = np.zeros((height, width), np.float32)
        check(lib().rr_primary_hits(self.h, ptr(cam), width, height, ptr(mesh), ptr(prim), ptr(dst)), "rr_primary_hits")
        return mesh, prim, dst

    def render_cost(self, cam, width, height, spp, bounces):
        """Path segments per pixel (instrumented kernel): the cost map of the cost-ordered queue experiment."""
        cam = np.ascontiguousarray(cam, CAMERA)
        cost = np.zeros((height, width), np.uint32)
        check(lib().rr_render_cost(self.h, ptr(cam), width, height, spp, bounces, ptr(cost)), "rr_render_cost")
        return cost

    def set_tile_order(self, tiles=None):
        """The following renders hand out exactly these tiles (row-major tile numbers), in this order; None: row-major."""
        if tiles is None or len(tiles) == 0:
            check(lib().rr_set_tile_order(self.h, None, 0), "rr_set_tile_order")
            return
        t = np.ascontiguousarray(tiles, np.uint32)
        check(lib().rr_set_tile_order(self.h, ptr(t), len(t)), "rr_set_tile_order")

    def bvh(self, which=0):
        n = C.c_uint64()
        check(lib().rr_bvh_size(self.h, which, C.byref(n)), "rr_bvh_size")
        n = int(n.value)
        m = max(n, 1)
        out = dict(codes=np.zeros(m, np.uint64), order=np.zeros(m, np.uint32), left=np.zeros(m, np.int32),
                   right=np.zeros(m, np.int32), parent=np.zeros(m, np.int32), bounds=np.zeros((m, 6), np.float32))
        check(lib().rr_bvh_read(self.h, which, ptr(out["codes"]), ptr(out["order"]), ptr(out["left"]), ptr(out["right"]),
                                ptr(out["parent"]), ptr(out["bounds"])), "rr_bvh_read")
        return {k: v[:n] for k, v in out.items()}

    # -- multi-process tile queue -------------------------------------------------
    def queue_export(self, width, height):
        q = np.zeros(_abi_handle_bytes(), np.uint8)
        f = np.zeros(_abi_handle_bytes(), np.uint8)
        check(lib().rr_queue_export(self.h, width, height, ptr(q), ptr(f)), "rr_queue_export")
        return q, f

    def queue_import(self, width, height, q, f):
        q = np.ascontiguousarray(q, np.uint8)
        f = np.ascontiguousarray(f, np.uint8)
        check(lib().rr_queue_import(self.h, width, height, ptr(q), ptr(f)), "rr_queue_import")

    def queue_reset(self):
        check(lib().rr_queue_reset(self.h), "rr_queue_reset")

    def render_shared(self, cam, width, height, spp, bounces, frame_index=0, tile=0):
        cam = np.ascontiguousarray(cam, CAMERA)
        st = Stats()
        check(lib().rr_render_shared(self.h, ptr(cam), width, height, spp, bounces, frame_index, tile, C.byref(st)),
              "rr_render_shared")
        return st.as_dict()

    def render_strided(self, cam, width, height, spp, bounces, rank, world, frame_index=0, tile=0):
        cam = np.ascontiguousarray(cam, CAMERA)
        st = Stats()
        check(lib().rr_render_strided(self.h, ptr(cam), width, height, spp, bounces, frame_index, tile, rank, world,
                                      C.byref(st)), "rr_render_strided")
        return st.as_dict()

    def frame_device_ptr(self):
        p, b = C.c_uint64(), C.c_uint64()
        check(lib().rr_frame_device_ptr(self.h, C.byref(p), C.byref(b)), "rr_frame_device_ptr")
        return int(p.value), int(b.value)


def _abi_handle_bytes() -> int:
    return 64  # RR_IPC_HANDLE_BYTES
