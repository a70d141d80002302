"""Scan-directory ingest: the package's `RTAB_utils.ios_rtab` readers against vectors produced by the
reference's own `RTAB2Cache` readers (tests/golden/make_golden_ingest.py), and the streamed fusion from a directory
against the oracle."""
import importlib

import cv2
import numpy as np
import pytest

from conftest import PKG_NAME, load_golden, small_scene
from oracle import f3d_oracle as orc


def _write_export(root, pose_text, calib_text, depths, ids):
    (root / "depth").mkdir(parents=True)
    (root / "poses.txt").write_bytes(bytes(pose_text))
    (root / "calibration.yaml").write_bytes(bytes(calib_text))
    for d, i in zip(depths, ids):
        assert cv2.imwrite(str(root / "depth" / f"{int(i)}.png"), d)


def test_readers_match_reference(tmp_path):
    rtab = importlib.import_module(PKG_NAME + ".RTAB_utils.ios_rtab")
    g = load_golden("g6_ingest")
    _write_export(tmp_path, g["pose_text"], g["calibration_text"], g["depths_u16"], g["ids"])
    K = rtab.read_intrinsic(tmp_path / "calibration.yaml")
    assert K.shape == (3, 3) and np.array_equal(K, g["intrinsic"])
    assert np.array_equal(rtab.resize_camera_matrix(K, 64 / 720, 48 / 960), g["scaled_64x48_from_720x960"])
    assert np.array_equal(rtab.resize_camera_matrix(K, 1920 / 720, 1440 / 960), g["scaled_1920x1440_from_720x960"])
    for tag in ("all", "slice_pad"):
        a, b, pad = [int(v) for v in g[f"{tag}_range"]]
        a, b = (None if a < 0 else a), (None if b < 0 else b)
        cache = rtab.RTAB2Cache(str(tmp_path), str(tmp_path / "rgb"), str(tmp_path / "depth"), str(tmp_path / "poses.txt"),
                                a, b, 1, False, bool(pad))
        assert np.array_equal(cache.img_idx, g[f"{tag}_img_idx"]) and np.array_equal(cache.odo_xyz, g[f"{tag}_odo_xyz"])
        assert np.array_equal(cache.odo_wxyz, g[f"{tag}_odo_xyzw"]) and np.array_equal(cache.odo_timestamp, g[f"{tag}_stamp"])
        for k in range(len(cache.img_idx)):
            d = rtab.read_depth_png(cache.depth_file(k), cache.padding)
            assert d.dtype == np.uint16 and np.array_equal(d.astype(np.float64), g[f"{tag}_depths"][k])
        rts = cache.rts((960, 720, 3), (48, 64))
        assert np.array_equal(rts["intrinsicScaled"], g["scaled_64x48_from_720x960"])
        assert rts["odo_wxyz"] is cache.odo_wxyz and tuple(rts["Depth_res"]) == (48, 64)
    with pytest.raises(ValueError):
        cv2.imwrite(str(tmp_path / "bad.png"), np.zeros((4, 4), np.uint8))
        rtab.read_depth_png(tmp_path / "bad.png")


@pytest.mark.gpu
def test_label_scan_directory_vs_oracle(engine, scenes, tmp_path):
    """Pose text + calibration.yaml + depth / mask PNGs on disk -> streamed decode, H2D, GPU mask resize, fused votes and
    labels; compared with the oracle on the arrays the files hold."""
    rtab = importlib.import_module(PKG_NAME + ".RTAB_utils.ios_rtab")
    g = load_golden("g6_ingest")
    s = small_scene(scenes, orc, npoints=20011, nframes=7, width=160, height=120, seed=83, border=0, block=16)
    F, H, W = 7, s["H"], s["W"]
    ids = [5, 6, 9, 10, 11, 15, 21]
    xyzw = s["wxyz"][:, [1, 2, 3, 0]]
    text = "".join("%.6f %.6f %.6f %.6f %.6f %.6f %.6f %.6f %d\n" % (1000.0 + f, *s["t"][f], *xyzw[f], ids[f]) for f in range(F))
    depth = s["depths"].copy()
    depth[:, 30:40, 50:70] = 1234                                   # make the 10-pixel border zeroing matter
    depth[:, :5, :] = 900
    _write_export(tmp_path, text.encode(), g["calibration_text"], depth, ids)
    (tmp_path / "masks").mkdir()
    big = scenes.block_masks((2 * H, 2 * W), F, seed=85, block=24)   # masks live at RGB resolution (voting.py:93)
    for f in range(F):
        if f != 3:                                                   # one frame without a mask: skipped like voting.py:45-54
            assert cv2.imwrite(str(tmp_path / "masks" / f"{ids[f]}.png"), big[f])
    cache = rtab.RTAB2Cache(str(tmp_path), str(tmp_path / "rgb"), str(tmp_path / "depth"), str(tmp_path / "poses.txt"),
                            None, None, 1, False, True)
    assert np.array_equal(cache.rts((960, 720, 3), (H, W))["intrinsicScaled"], s["K"])
    votes, classes, fl = rtab.label_scan(s["points"], cache, tmp_path / "masks", (960, 720, 3), (0.1, 4.0), 0.05, 133, 0.5,
                                         [86, 114, 115], chunk_frames=3, workers=4, return_labeler=True)
    keep = [f for f in range(F) if f != 3]
    pad = depth[keep].copy()
    pad[:, :10, :] = 0
    pad[:, -10:, :] = 0
    pad[:, :, :10] = 0
    pad[:, :, -10:] = 0
    small = np.stack([cv2.resize(big[f], (W, H), interpolation=cv2.INTER_NEAREST) for f in keep])
    ov = orc.fuse_project_vote(s["points"], s["K"], W, H, s["wxyz"][keep], s["t"][keep], pad, small, 134, 0, 0.05, 0.1, 4.0, 4.0)
    assert votes.dtype == np.float64 and np.array_equal(votes, ov.astype(np.float64)) and ov.sum() > 0
    assert classes.dtype == np.int64 and np.array_equal(classes, orc.segment(ov, 133, 0.5, [86, 114, 115]))
    assert fl.nframes == len(keep)
