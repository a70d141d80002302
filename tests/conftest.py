import importlib
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

PKG_NAME = "3d-point-cloud-segmentation-using-2d-img-segmentation_b200"
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module(PKG_NAME)


@pytest.fixture(scope="session")
def scenes():
    return importlib.import_module(PKG_NAME + ".scenes")


@pytest.fixture(scope="session")
def engine(pkg):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return importlib.import_module(PKG_NAME + ".engine")


def load_golden(name):
    return dict(np.load(GOLDEN / f"{name}.npz"))


def small_scene(scenes, orc, npoints=30000, nframes=6, width=160, height=120, seed=7, zmax=4.0, border=4, block=8,
                base="C1"):
    """Seeded small scene with oracle-rendered depth; shared by CPU and GPU tests."""
    spec = scenes.scaled_spec(base, npoints=npoints, nframes=nframes, width=width, height=height, seed=seed)
    K = scenes.scaled_intrinsics(width, height)
    wxyz, t = scenes.make_poses(spec)
    pts = scenes.make_cloud(spec)
    p64 = pts.astype(np.float64)
    eyes, look, nrm = orc.frustum_data(K, width, height, wxyz, t)
    depths = np.stack([orc.zero_border(orc.zbuffer_splat(p64, K, width, height, wxyz[f], t[f], eyes[f], look[f], nrm[f],
                                                         zmax), border) for f in range(nframes)])
    masks = scenes.block_masks((height, width), nframes, seed=seed, block=block)
    return dict(points=pts, K=K, W=width, H=height, wxyz=wxyz, t=t, depths=depths, masks=masks, zmax=zmax)
