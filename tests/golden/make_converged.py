"""Generates tests/golden/converged_ref.npz: the reference's CONVERGED image of its default scene (north_star: "PSNR >= 50 dB
against the reference's converged image"; SURVEY.md 8c, level L4 of the parity ladder).

Run in the build container (needs /root/reference and oracle/_ref built by `make -C oracle ref`; about 25 minutes on
8 cores, 9 of them the brute-force pass):   python tests/golden/make_converged.py

The default scene (src/main.cpp:246-304, a 16 x 8 UV sphere standing in for the unshipped knight.obj) is rendered at
32 x 32 pixels, 2^18 samples per pixel, 50 bounces by BOTH host builds of the reference's kernel text (oracle/ref_shim):
  * `fast`   -- -O3 -ffast-math, the analogue of the -cl-fast-relaxed-math build the reference JITs (src/image.hpp:49):
                the "reference's converged image" of the criterion;
  * `strict` -- IEEE arithmetic without contraction, the numerics contract the CUDA kernel is bit-exact against.
and by the C restatement (oracle/rr_oracle.c) in the two ways that DEFINE the result of this repo:
  * brute force -- every primitive of every mesh tested for every path segment (229 G triangle tests), no hierarchy;
  * its walk of the LBVH with delta-inflated boxes -- must be the same bits (the hierarchy only culls), asserted here.
Stored: the upload arrays, the camera, the float radiance images (mean incoming light per pixel, before the clamp and
gamma of src/Trace.cl:646-650) `rad_fast`, `rad_strict`, `rad_definition`, and the path-segment counts.

What the run of 2026-10 found (recorded in DESIGN.md 3): brute force == LBVH walk, 962 202 654 path segments; the compiled
strict reference, which walks its own SAH hierarchy with exact (not inflated) boxes and first-found-wins on equal
distances, differs from its own brute force in 13 of the 1 024 pixels (962 202 643 segments, largest radiance difference
9.8e-5) -- and the restatement walking the reference's node list reproduces those 13 pixels bit for bit.
"""
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.pyoracle import Oracle, Reference, mesh_ranges_from_gpunodes  # noqa: E402
from ripoff_raytracer_b200 import scenes  # noqa: E402

OUT = Path(__file__).resolve().parent
W = H = 32
SPP = 1 << 18
BOUNCES = 50


def psnr(a, b):
    """PSNR of two radiance images on the displayable range: clamp to [0, 1] (src/Trace.cl:646), peak 1."""
    d = np.clip(a.astype(np.float64), 0, 1) - np.clip(b.astype(np.float64), 0, 1)
    mse = float(np.mean(d * d))
    return float("inf") if mse == 0 else 10.0 * np.log10(1.0 / mse)


def main():
    data = {}
    for variant in ("strict", "fast"):
        ref = Reference(variant)
        with tempfile.TemporaryDirectory() as td:
            obj = Path(td) / "knight.obj"
            v, n, f = scenes.uv_sphere(16, 8)
            scenes.write_obj(obj, v, n, f)
            tris, meshes, nodes = ref.scene_default(obj)
        cam = ref.default_camera(W, H)
        if variant == "strict":
            data.update(tris=tris, meshes=meshes, ranges=mesh_ranges_from_gpunodes(meshes, nodes), cam=cam, W=W, H=H, spp=SPP,
                        bounces=BOUNCES)
            strict_nodes = nodes  # the node list the strict build walked
        t0 = time.time()
        _, rad = ref.render(cam, W, H, SPP, BOUNCES, radiance=True)
        print(variant, f"{time.time() - t0:.1f} s", f"{W * H * SPP / (time.time() - t0) / 1e6:.2f} Msamples/s", flush=True)
        data[f"rad_{variant}"] = rad
    bits = lambda a: a.view(np.uint32)  # noqa: E731
    t0 = time.time()
    _, walk, st_walk = Oracle(data["tris"], data["meshes"], data["ranges"]).render(data["cam"], W, H, SPP, BOUNCES, radiance=True)
    print("restatement, LBVH walk", f"{time.time() - t0:.1f} s", st_walk, flush=True)
    t0 = time.time()
    _, brute, st_brute = Oracle(data["tris"], data["meshes"], data["ranges"]).brute_force(True).render(data["cam"], W, H, SPP, BOUNCES, radiance=True)
    print("restatement, brute force", f"{time.time() - t0:.1f} s", st_brute, flush=True)
    assert np.array_equal(bits(walk), bits(brute)) and st_walk["rays"] == st_brute["rays"], "the hierarchy must only cull"
    _, sah, st_sah = Oracle(data["tris"], data["meshes"], data["ranges"], ref_gpunodes=strict_nodes).render(data["cam"], W, H, SPP, BOUNCES, radiance=True)
    assert np.array_equal(bits(sah), bits(data["rad_strict"])), "restatement on the reference's node list == compiled reference"
    data["rad_definition"] = brute
    data["rays_definition"] = np.uint64(st_brute["rays"])
    data["rays_reference_walk"] = np.uint64(st_sah["rays"])
    differ = (bits(brute) != bits(data["rad_strict"])).any(axis=2)
    print("pixels where the compiled reference differs from brute force:", int(differ.sum()),
          "largest difference:", float(np.max(np.abs(brute - data["rad_strict"]))))
    print("PSNR strict vs fast:", psnr(data["rad_strict"], data["rad_fast"]),
          "max abs:", float(np.max(np.abs(np.clip(data["rad_strict"], 0, 1) - np.clip(data["rad_fast"], 0, 1)))))
    np.savez_compressed(OUT / "converged_ref.npz", **data)


if __name__ == "__main__":
    main()
