"""Deterministic synthetic inputs for the configs of BASELINE.json.

The reference ships no models (`*.obj` is git-ignored, its default
`knight.obj` is absent), so every input is generated here from a seed.  Meshes
are written in the OBJ dialect the reference's loader accepts
(src/readobj.hpp:289-344: `v`, `vn`, `f a//a` triangles, normals mandatory) and
sized like the default scene (10^2..10^3 units) so the EPSILON = 1e-6 logic of
the kernel behaves.
"""
from __future__ import annotations

import numpy as np

from ._abi import MATERIAL_SOLID, SPHERE, TRIANGLE


# ------------------------------------------------------------------ meshes --
def uv_sphere(nu: int = 48, nv: int = 24, radius: float = 100.0, center=(0.0, 100.0, 0.0)):
    """Closed UV sphere standing on y = 0 (stand-in for knight.obj): 2*nu*(nv-1) triangles."""
    verts = [(0.0, 1.0, 0.0)]
    for j in range(1, nv):
        th = np.pi * j / nv
        for i in range(nu):
            ph = 2.0 * np.pi * i / nu
            verts.append((np.sin(th) * np.cos(ph), np.cos(th), np.sin(th) * np.sin(ph)))
    verts.append((0.0, -1.0, 0.0))
    n = np.asarray(verts, dtype=np.float64)
    faces = []
    ring = lambda j, i: 1 + (j - 1) * nu + (i % nu)
    south = len(verts) - 1
    for i in range(nu):
        faces.append((0, ring(1, i + 1), ring(1, i)))
    for j in range(1, nv - 1):
        for i in range(nu):
            a, b, c, d = ring(j, i), ring(j, i + 1), ring(j + 1, i + 1), ring(j + 1, i)
            faces.append((a, b, c))
            faces.append((a, c, d))
    for i in range(nu):
        faces.append((south, ring(nv - 1, i), ring(nv - 1, i + 1)))
    v = (n * radius + np.asarray(center)).astype(np.float32)
    return v, n.astype(np.float32), np.asarray(faces, dtype=np.int64)


def _icosahedron():
    t = (1.0 + 5.0**0.5) / 2.0
    v = np.array(
        [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array(
        [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
         (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10),
         (8, 6, 7), (9, 8, 1)], dtype=np.int64)
    return v, f


def icosphere(subdiv: int):
    """Unit icosphere: 20 * 4**subdiv triangles, outward (CCW) winding."""
    v, f = _icosahedron()
    for _ in range(subdiv):
        nv0 = len(v)
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], axis=0)
        es = np.sort(e, axis=1)
        key = es[:, 0] * nv0 + es[:, 1]
        uniq, inv = np.unique(key, return_inverse=True)
        a, b = uniq // nv0, uniq % nv0
        mid = v[a] + v[b]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        v = np.concatenate([v, mid], axis=0)
        nf = len(f)
        m01, m12, m20 = nv0 + inv[:nf], nv0 + inv[nf:2 * nf], nv0 + inv[2 * nf:]
        f = np.concatenate(
            [np.stack([f[:, 0], m01, m20], 1), np.stack([f[:, 1], m12, m01], 1), np.stack([f[:, 2], m20, m12], 1),
             np.stack([m01, m12, m20], 1)], axis=0)
    return v, f


def _fbm(p: np.ndarray, seed: int, octaves: int = 5) -> np.ndarray:
    """Smooth pseudo-noise on points p (n,3): sum of random-phase sinusoids."""
    rng = np.random.default_rng(seed)
    out = np.zeros(len(p))
    amp, freq = 1.0, 1.5
    for _ in range(octaves):
        for _k in range(4):
            d = rng.normal(size=3)
            d /= np.linalg.norm(d)
            out += amp * np.sin(freq * (p @ d) + rng.uniform(0, 2 * np.pi)) / 4.0
        amp *= 0.5
        freq *= 2.1
    return out


def vertex_normals(v: np.ndarray, f: np.ndarray) -> np.ndarray:
    fn = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
    n = np.zeros_like(v, dtype=np.float64)
    for k in range(3):
        np.add.at(n, f[:, k], fn)
    ln = np.linalg.norm(n, axis=1, keepdims=True)
    ln[ln == 0] = 1.0
    return n / ln


def displaced_icosphere(subdiv: int, radius: float = 100.0, center=(0.0, 100.0, 0.0), amplitude: float = 0.12,
                        seed: int = 3):
    """Closed bumpy blob with smooth vertex normals; 20*4**subdiv triangles (subdiv 6 = 81 920)."""
    v, f = icosphere(subdiv)
    r = 1.0 + amplitude * _fbm(v, seed)
    v = v * r[:, None]
    n = vertex_normals(v, f)
    vv = (v * radius + np.asarray(center)).astype(np.float32)
    return vv, n.astype(np.float32), f


def heightfield(n: int, size: float = 400.0, height: float = 60.0, base: float = 40.0, seed: int = 4):
    """n x n quads -> 2 n^2 triangles, an open terrain sheet facing +y."""
    g = np.linspace(-0.5, 0.5, n + 1)
    X, Z = np.meshgrid(g, g, indexing="xy")
    p = np.stack([X.ravel() * 3.0, np.zeros(X.size), Z.ravel() * 3.0], 1)
    y = base + height * (0.5 + 0.5 * _fbm(p, seed, octaves=6))
    v = np.stack([X.ravel() * size, y, Z.ravel() * size], 1)
    idx = np.arange((n + 1) * (n + 1)).reshape(n + 1, n + 1)
    a, b, c, d = idx[:-1, :-1].ravel(), idx[:-1, 1:].ravel(), idx[1:, 1:].ravel(), idx[1:, :-1].ravel()
    # (a, d, c), (a, c, b): normal points to +y
    f = np.concatenate([np.stack([a, d, c], 1), np.stack([a, c, b], 1)], axis=0)
    nrm = vertex_normals(v, f)
    return v.astype(np.float32), nrm.astype(np.float32), f.astype(np.int64)


def merge_meshes(parts):
    """Concatenate (v, n, f) meshes into one."""
    vs, ns, fs, off = [], [], [], 0
    for v, n, f in parts:
        vs.append(v)
        ns.append(n)
        fs.append(f + off)
        off += len(v)
    return np.concatenate(vs), np.concatenate(ns), np.concatenate(fs)


def blob_cluster(count: int, subdiv: int, seed: int, extent=(300.0, 120.0, 300.0), radius=(14.0, 30.0)):
    """`count` displaced icospheres scattered above y = 0, merged into one mesh."""
    rng = np.random.default_rng(seed)
    parts = []
    for k in range(count):
        r = rng.uniform(*radius)
        c = (rng.uniform(-extent[0], extent[0]), r * 0.9 + rng.uniform(0, extent[1]), rng.uniform(-extent[2], extent[2]))
        parts.append(displaced_icosphere(subdiv, radius=r, center=c, seed=seed * 1000 + k))
    return merge_meshes(parts)


def write_obj(path, v: np.ndarray, n: np.ndarray, f: np.ndarray) -> None:
    """`v` / `vn` / `f a//a` with %.9g (float32 round-trips exactly)."""
    with open(path, "w") as fh:
        fh.write("# generated by ripoff_raytracer_b200.scenes\n")
        np.savetxt(fh, v, fmt="v %.9g %.9g %.9g")
        np.savetxt(fh, n, fmt="vn %.9g %.9g %.9g")
        f1 = f + 1
        np.savetxt(fh, np.stack([f1[:, 0], f1[:, 0], f1[:, 1], f1[:, 1], f1[:, 2], f1[:, 2]], 1),
                   fmt="f %d//%d %d//%d %d//%d")


def mesh_triangles(v: np.ndarray, n: np.ndarray, f: np.ndarray) -> np.ndarray:
    """The Triangle array the OBJ loader would produce for this mesh (src/readobj.hpp:333-342)."""
    t = np.zeros(len(f), dtype=TRIANGLE)
    for k, (pn, nn) in enumerate((("posA", "normalA"), ("posB", "normalB"), ("posC", "normalC"))):
        t[pn][:, :3] = v[f[:, k]]
        t[nn][:, :3] = n[f[:, k]]
    return t


# ----------------------------------------------------------------- spheres --
def random_spheres(count: int = 1024, seed: int = 2, box=(600.0, 300.0, 600.0), radius=(4.0, 16.0),
                   emissive_fraction: float = 0.10) -> np.ndarray:
    """Non-overlapping spheres above y = 0 (config C2): 90 % Solid albedo U(0.2,0.9)^3, 10 % emissive x4."""
    rng = np.random.default_rng(seed)
    out = np.zeros(count, dtype=SPHERE)
    centers = np.zeros((count, 3))
    radii = np.zeros(count)
    k = 0
    cell = 2.0 * radius[1]
    grid: dict = {}
    while k < count:
        r = rng.uniform(*radius)
        c = np.array([rng.uniform(-box[0] / 2, box[0] / 2), r + rng.uniform(0, box[1] - 2 * r),
                      rng.uniform(-box[2] / 2, box[2] / 2)])
        key = tuple((c // cell).astype(int))
        ok = True
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dz in (-1, 0, 1):
                    for j in grid.get((key[0] + dx, key[1] + dy, key[2] + dz), ()):
                        if np.linalg.norm(centers[j] - c) < radii[j] + r:
                            ok = False
        if not ok:
            continue
        centers[k], radii[k] = c, r
        grid.setdefault(key, []).append(k)
        k += 1
    out["center"][:, :3] = centers.astype(np.float32)
    out["radius"] = radii.astype(np.float32)
    emissive = rng.uniform(size=count) < emissive_fraction
    albedo = rng.uniform(0.2, 0.9, size=(count, 3)).astype(np.float32)
    out["material"]["type"] = MATERIAL_SOLID
    out["material"]["ior"] = 1.0
    out["material"]["color"][:, :3] = albedo
    out["material"]["emissionColor"][:, :3] = np.where(emissive[:, None], 1.0, 0.0)
    out["material"]["emissionStrength"] = np.where(emissive, 4.0, 0.0)
    out["material"]["reflectiveness"] = 0.0
    out["material"]["specularProbability"] = 0.0
    return out
