import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(Path(__file__).resolve().parent))  # tests/cases.py
GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def golden_small():
    return dict(np.load(GOLDEN / "default_small.npz"))


@pytest.fixture(scope="session")
def golden_wide():
    return dict(np.load(GOLDEN / "default_wide.npz"))


@pytest.fixture(scope="session")
def golden_zoo():
    """Checker / Glassy / OneSided / specular meshes with pitch, yaw, roll and scale, rendered by the compiled
    reference with its own SAH hierarchy (tests/golden/make_golden.py zoo)."""
    return dict(np.load(GOLDEN / "zoo_ref.npz"))


@pytest.fixture(scope="session")
def golden_converged():
    """The reference's default scene at 32 x 32, 2^18 spp, 50 bounces: radiance of its -ffast-math and of its strict
    build (tests/golden/make_converged.py)."""
    return dict(np.load(GOLDEN / "converged_ref.npz"))


def psnr_radiance(a, b):
    """PSNR of two float radiance images on the displayable range: clamped to [0, 1] (src/Trace.cl:646), peak 1."""
    d = np.clip(a.astype(np.float64), 0, 1) - np.clip(b.astype(np.float64), 0, 1)
    mse = float(np.mean(d * d))
    return float("inf") if mse == 0 else 10.0 * np.log10(1.0 / mse)


@pytest.fixture(scope="session")
def golden_rng():
    return dict(np.load(GOLDEN / "rng.npz"))


@pytest.fixture(scope="session")
def knight_obj(tmp_path_factory):
    """Stand-in for the reference's unshipped knight.obj: a 2 208-triangle UV sphere."""
    from ripoff_raytracer_b200 import scenes

    p = tmp_path_factory.mktemp("obj") / "knight.obj"
    scenes.write_obj(p, *scenes.uv_sphere(48, 24))
    return p


@pytest.fixture(scope="session")
def renderer():
    import ripoff_raytracer_b200 as rr

    r = rr.Renderer()
    yield r
    r.close()


def psnr(a, b, peak=255.0):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return float("inf") if mse == 0 else 10.0 * np.log10(peak * peak / mse)


@pytest.fixture(scope="session")
def golden_video():
    """Progressive averages and video poses produced by the reference (tests/golden/make_golden.py progressive_video)."""
    return dict(np.load(GOLDEN / "progressive_video.npz"))
