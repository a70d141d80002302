"""process3D / FrameData mirror: the cache loader against vectors of the UNMODIFIED reference `FrameData` (CPU), and
`process3DSeg` on a given cloud against the oracle (GPU)."""
import importlib
import pickle
import sys
from pathlib import Path

import numpy as np
import pytest

from conftest import PKG_NAME, load_golden, small_scene
from oracle import f3d_oracle as orc

sys.path.insert(0, str(Path(__file__).parent / "golden"))
from make_golden_framedata import write_cache  # noqa: E402  (pure pickle writer; the reference import in that module is lazy-safe)


def test_framedata_matches_reference_loader(tmp_path):
    fus = importlib.import_module(PKG_NAME + ".Fusion3DSeg.fusion")
    g = load_golden("g8_framedata")
    tof = write_cache(tmp_path, g["depth_mm"], g["frame_numbers"])
    F, h, w = g["depth_mm"].shape
    for dec in (1, 2, 3):
        fd = fus.FrameData(tof, (0.1, 4), dec, (h, w))
        assert len(fd) == F
        for i in range(F):
            name, pts, nrm, clr, valid = fd[i]
            assert name == str(g["names"][i]) and pts.shape == (h * w, 3)
            assert np.array_equal(valid, g[f"valid_dec{dec}"][i])
            name2, d = fd.depth_mm(i)
            assert name2 == name and d.dtype == np.uint16
            assert np.array_equal(d != 0, (g[f"valid_dec{dec}"][i] & (g["depth_mm"][i].reshape(-1) != 0)).reshape(h, w))
            assert np.array_equal(d[d != 0], g["depth_mm"][i][d != 0])


@pytest.mark.gpu
def test_process3dseg_on_given_cloud(engine, scenes, tmp_path):
    p3d = importlib.import_module(PKG_NAME + ".Fusion3DSeg.process3D")
    s = small_scene(scenes, orc, npoints=12000, nframes=5, width=96, height=72, seed=21)
    F, h, w = s["depths"].shape
    frames = np.array([4, 7, 8, 15, 16])
    write_cache(tmp_path, s["depths"], frames)
    xyzw = s["wxyz"][:, [1, 2, 3, 0]]
    rts = {"intrinsic": s["K"], "intrinsicScaled": s["K"], "odo_wxyz": xyzw, "odo_xyz": s["t"], "RGB_res": (h, w, 3), "Depth_res": (h, w)}
    with open(tmp_path / "PointcloudMergeResults" / "rtscameradata_x.pkl", "wb") as fp:
        pickle.dump(rts, fp)
    out = tmp_path / "out"
    pts, norms, clrs, nmerges, occ, nframes, hw, adj = p3d.process3DSeg(str(tmp_path), str(out), radius=0.05, point_range=(0.1, 4),
                                                                        decimation=1, cloud=s["points"])
    assert nframes == F and tuple(hw) == (h, w) and np.array_equal(pts, s["points"].astype(np.float64))
    eyes, look, nrm = orc.frustum_data(s["K"], w, h, s["wxyz"], s["t"])
    p64 = s["points"].astype(np.float64)
    tot = np.zeros(len(p64), np.int64)
    seen = np.zeros(len(p64), np.int64)
    for f in range(F):
        want = orc.frame_uv2pt(p64, s["K"], w, h, s["wxyz"][f], s["t"][f], eyes[f], look[f], nrm[f], s["depths"][f], 0, 0.05, 0.1, 4.0, 4.0)
        got = np.load(out / "fusion" / "uv2pt" / f"{frames[f]}.npy")
        assert got.dtype == np.int32 and np.array_equal(got, want)
        hit = want[want >= 0]
        tot += np.bincount(hit, minlength=len(p64))
        seen[np.unique(hit)] += 1
    assert np.array_equal(nmerges, tot) and np.array_equal(occ, seen)
    oip, oix = orc.radius_adjacency(p64, 0.1)                                     # KDTree.query_radius(points, r=2*radius), fusion.py:374-375
    assert len(adj) == len(p64) and all(np.array_equal(adj[i], oix[oip[i]:oip[i + 1]]) for i in range(0, len(p64), 97))
    # decimation: only the ::2 lattice can be matched
    p3d.process3DSeg(str(tmp_path), str(out), radius=0.05, point_range=(0.1, 4), decimation=2)       # cloud from fusion_data.pkl
    got = np.load(out / "fusion" / "uv2pt" / f"{frames[0]}.npy").reshape(h, w)
    lattice = np.zeros((h, w), bool)
    lattice[::2, ::2] = True
    assert (got[~lattice] == -1).all() and (got[lattice] >= 0).any()
