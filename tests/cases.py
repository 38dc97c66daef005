"""Scene builders shared by the CPU (oracle) and GPU parity tests."""
import numpy as np

from ripoff_raytracer_b200 import _abi, scenes


def far_origin_case(ratio, scale=1.0, W=192, H=128):
    """One rotated 5 120-triangle mesh seen by a camera `ratio` mesh extents away.  The rounding
    error of the slab test grows with |origin|, so this is where a hierarchy that "only culls" can start to lose hits;
    `scale` < 1 moves the mesh-local origin out further still (origin / scale, src/Trace.cl:127-130)."""
    v, n, f = scenes.displaced_icosphere(4, radius=40.0, center=(0.0, 0.0, 0.0), seed=3)
    t = scenes.mesh_triangles(v, n, f)
    m = np.zeros(1, _abi.MESH)
    m["scale"] = scale
    m["pitch"], m["yaw"], m["roll"] = 0.3, 0.7, -0.2
    mm = m["material"]
    mm["color"][:, :3] = 0.8
    mm["emissionColor"][:, :3] = (1.0, 0.7, 0.4)
    mm["emissionStrength"] = 1.0
    r = np.zeros(1, _abi.MESH_RANGE)
    r["numTriangles"] = len(t)
    ext = 40.0 * scale
    d = np.array([0.3, 0.5, 0.81])
    d /= np.linalg.norm(d)
    cam = np.zeros(1, _abi.CAMERA)
    cam["position"][0, :3] = (d * ratio * ext).astype(np.float32)
    cam["yaw"] = np.arctan2(-d[0], -d[2])
    cam["pitch"] = np.arcsin(d[1])
    cam["fov"] = np.degrees(2 * np.arctan(1.3 / ratio))
    cam["aspectRatio"] = np.float32(W) / np.float32(H)
    return t, m, r, cam, W, H
