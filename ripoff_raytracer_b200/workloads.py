"""The configurations of BASELINE.json as concrete synthetic inputs.

Every workload is (scene arrays, camera, width, height, spp, bounces).  Meshes go through the
reference's own scene assembly path (one OBJ-style mesh + addCornellBoxToScene around it, mesh
appended last) so the inputs are what a user of the reference would hand to generateBuffers.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _abi, scenes
from .api import Scene, default_camera


@dataclass
class Workload:
    name: str
    description: str
    scene: Scene
    cam: np.ndarray
    width: int
    height: int
    spp: int
    bounces: int

    @property
    def samples(self) -> int:
        return self.width * self.height * self.spp


def _white_mesh(scale=1.0, yaw=0.0):
    m = np.zeros(1, _abi.MESH)
    m["scale"] = scale
    m["yaw"] = yaw
    mm = m["material"]
    mm["type"] = _abi.MATERIAL_SOLID  # src/main.cpp:256-264
    mm["ior"] = 1.0
    mm["color"][:, :3] = 1.0
    mm["specularProbability"] = 1.0
    return m


def _cornell_around(tris, scale=1.0, yaw=0.0, spheres=None) -> Scene:
    """mesh + addCornellBoxToScene(mesh) + meshList.emplace_back(mesh) (src/main.cpp:246-298)."""
    s = Scene()
    rng = s.add_triangles(tris)
    mesh = _white_mesh(scale, yaw)
    s.add_cornell(mesh, rng)
    s.add_mesh(mesh, rng)
    if spheres is not None:
        s.add_spheres(spheres)
    return s


def _camera(width, height, pos, pitch=0.0, yaw=3.14, fov=90.0):
    cam = default_camera(width, height)
    cam["position"][0, :3] = pos
    cam["pitch"] = pitch
    cam["yaw"] = yaw
    cam["fov"] = fov
    return cam


def c1_default(width=512, height=512, spp=50, bounces=50) -> Workload:
    """configs[0]: the reference's built-in scene at its default resolution/spp (src/settings.hpp:34-43) with a
    2 208-triangle UV sphere standing in for the unshipped knight.obj."""
    v, n, f = scenes.uv_sphere(48, 24)
    s = _cornell_around(scenes.mesh_triangles(v, n, f), scale=0.5, yaw=5.5)
    return Workload("c1_default", "default scene (UV-sphere stand-in for knight.obj, 2222 tris, 8 meshes)", s,
                    default_camera(width, height), width, height, spp, bounces)


def c2_spheres(width=1920, height=1080, spp=64, bounces=8, count=1024) -> Workload:
    """configs[1]: spheres only (extension: the reference kernel has no sphere primitive) + floor/ceiling quads."""
    s = Scene()
    s.add_quad((-400, 0, -400), (400, 0, -400), (400, 0, 400), (-400, 0, 400), (0, 1, 0), (0.5, 0.5, 0.5))
    s.add_quad((-400, 320, -400), (400, 320, -400), (400, 320, 400), (-400, 320, 400), (0, -1, 0), (0.9, 0.9, 0.9))
    s.add_spheres(scenes.random_spheres(count, seed=2))
    cam = _camera(width, height, (0.0, 170.0, 560.0), pitch=0.08, fov=70.0)
    return Workload("c2_spheres", f"{count} spheres + floor/ceiling quads", s, cam, width, height, spp, bounces)


def c3_mesh100k(width=1920, height=1080, spp=256, bounces=50) -> Workload:
    """configs[2]: one procedural OBJ-style mesh, 81 920 triangles (displaced icosphere, subdivision 6)."""
    v, n, f = scenes.displaced_icosphere(6, seed=3)
    s = _cornell_around(scenes.mesh_triangles(v, n, f), scale=0.5, yaw=5.5)
    return Workload("c3_mesh100k", "displaced icosphere 81920 tris in the Cornell box", s, default_camera(width, height),
                    width, height, spp, bounces)


def c4_mixed1m(width=3840, height=2160, spp=256, bounces=50, terrain=500, blobs=25, blob_subdiv=5, n_spheres=256,
               seed=4) -> Workload:
    """configs[3]: ~1 M triangles (terrain height-field 2*terrain^2 + `blobs` displaced icospheres of
    20*4^blob_subdiv) as ONE mesh + 256 spheres, inside the reference's Cornell box."""
    parts = [scenes.heightfield(terrain, size=800.0, height=70.0, base=0.0, seed=seed),
             scenes.blob_cluster(blobs, blob_subdiv, seed=seed, extent=(330.0, 160.0, 330.0), radius=(18.0, 42.0))]
    v, n, f = scenes.merge_meshes(parts)
    v[:, 1] += 90.0  # blobs float above the terrain, which spans y in [0, 70]
    v[: (terrain + 1) ** 2, 1] -= 90.0
    sph = scenes.random_spheres(n_spheres, seed=seed, box=(700.0, 200.0, 700.0), radius=(5.0, 14.0))
    sph["center"][:, 1] += 80.0
    s = _cornell_around(scenes.mesh_triangles(v, n, f), spheres=sph)
    # a brighter ceiling light than the default 100x100 quad would give this 1000-unit room
    light = s.mesh(6)["material"]
    light["emissionStrength"] = 60.0
    cam = _camera(width, height, (0.0, 260.0, 470.0), pitch=0.35, fov=90.0)
    ntri = len(f)
    return Workload("c4_mixed1m", f"{ntri} triangles (terrain + {blobs} blobs, one mesh) + {n_spheres} spheres in the Cornell box",
                    s, cam, width, height, spp, bounces)


def c5_mesh10m(width=7680, height=4320, spp=1024, bounces=50, seed=5) -> Workload:
    """configs[4]: ~10 M triangles (2 000^2 height-field + 100 subdivision-6 blobs)."""
    w = c4_mixed1m(width, height, spp, bounces, terrain=1000, blobs=100, blob_subdiv=6, n_spheres=256, seed=seed)
    w.name = "c5_mesh10m"
    return w


def instances(width=1920, height=1080, spp=8, bounces=16, count=1024, subdiv=3, seed=9) -> Workload:
    """many instances of ONE mesh (a displaced icosphere of 20*4^subdiv triangles, one triangle
    range shared by every MeshInfo -- the reference's meshList allows that, src/readobj.hpp:75-81) with random pose,
    scale and material, in a room of six walls and a ceiling light.  Exercises the top level over the mesh boxes."""
    rng = np.random.default_rng(seed)
    s = Scene()
    v, n, f = scenes.displaced_icosphere(subdiv, radius=10.0, center=(0.0, 0.0, 0.0), seed=seed)
    blob = s.add_triangles(scenes.mesh_triangles(v, n, f))
    side = 40.0 * max(count, 8) ** (1.0 / 3.0) + 60.0  # room edge grows with the instance count (constant density)
    h = 0.6 * side
    s.add_quad((-side, 0, -side), (side, 0, -side), (side, 0, side), (-side, 0, side), (0, 1, 0), (0.6, 0.6, 0.6))
    s.add_quad((-side, h, -side), (side, h, -side), (side, h, side), (-side, h, side), (0, -1, 0), (0.9, 0.9, 0.9))
    s.add_quad((-side, 0, -side), (side, 0, -side), (side, h, -side), (-side, h, -side), (0, 0, 1), (0.2, 0.7, 0.2))
    s.add_quad((-side, 0, -side), (-side, 0, side), (-side, h, side), (-side, h, -side), (1, 0, 0), (0.2, 0.2, 0.9))
    s.add_quad((side, 0, -side), (side, 0, side), (side, h, side), (side, h, -side), (-1, 0, 0), (0.9, 0.2, 0.2))
    s.add_quad((-0.5 * side, h - 1, -0.5 * side), (0.5 * side, h - 1, -0.5 * side), (0.5 * side, h - 1, 0.5 * side),
               (-0.5 * side, h - 1, 0.5 * side), (0, -1, 0), (1, 1, 1))
    lm = s.mesh(s.n_meshes - 1)["material"]
    lm["emissionColor"][0, :3] = 1.0
    lm["emissionStrength"] = 6.0
    m = np.zeros(count, _abi.MESH)
    m["pos"][:, :3] = rng.uniform((-0.9 * side, 12.0, -0.9 * side), (0.9 * side, 0.9 * h, 0.9 * side), size=(count, 3))
    m["pitch"], m["yaw"], m["roll"] = rng.uniform(-3.1, 3.1, size=(3, count)).astype(np.float32)
    m["scale"] = rng.choice(np.array([0.5, 0.8, 1.0, 1.3, 2.0], np.float32), size=count)
    mm = m["material"]
    mm["type"] = rng.choice(np.array([_abi.MATERIAL_SOLID] * 6 + [_abi.MATERIAL_GLASSY, _abi.MATERIAL_ONESIDED]), size=count)
    mm["ior"] = 1.45
    mm["color"][:, :3] = rng.uniform(0.25, 0.95, size=(count, 3))
    mm["specularProbability"] = rng.uniform(0.0, 0.6, size=count)
    mm["reflectiveness"] = rng.uniform(0.0, 0.9, size=count)
    for k in range(count):
        s.add_mesh(m[k:k + 1], blob)
    cam = _camera(width, height, (0.0, 0.55 * h, 0.95 * side), pitch=0.12, yaw=3.14159, fov=80.0)
    return Workload(f"instances_{count}", f"{count} instances of a {len(f)}-triangle mesh + 6 wall quads", s, cam, width, height,
                    spp, bounces)


WORKLOADS = {"instances": instances, "c1": c1_default, "c2": c2_spheres, "c3": c3_mesh100k, "c4": c4_mixed1m, "c5": c5_mesh10m}
