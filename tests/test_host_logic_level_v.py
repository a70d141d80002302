"""Level V drop-in (`get3DSeg.segment` / `remove_classes` -> `VotingSegmentation`) WITHOUT a device: the four GPU operators the
mirror composes are replaced by numpy stand-ins of the same contract (built on the oracle), so that what runs here is the
mirror's host logic -- file pairing by stem, frame batching, mask resize decision, the float64 `votes` view, the
`nclasses = 134` reload quirk (`voting.py:40`), the files `segment` writes -- against the vectors the unmodified reference
produced (g3).  The operators themselves are checked on the GPU (tests/test_gpu_parity.py, tests/test_gpu_dropin.py)."""
import importlib
import pickle

import numpy as np
import pytest
import torch

from conftest import PKG_NAME, load_golden
from oracle import f3d_oracle as orc


@pytest.fixture
def level_v(monkeypatch):
    voting = importlib.import_module(PKG_NAME + ".Fusion3DSeg.segUtils.voting")
    eng = voting.engine

    def resize_nearest(masks, height, width):
        return torch.as_tensor(np.stack([orc.resize_nearest(m, width, height) for m in masks.numpy()]))

    def vote_uv2pt(votes_packed, uv2pt, mask, first_tag):
        # contract at the vote_finalize boundary: one count per distinct (point, class) pair of each frame
        v = votes_packed.numpy()
        for u, m in zip(uv2pt.numpy(), mask.numpy()):
            orc.vote_uv2pt(v, u, m)
        return votes_packed

    def resolve_labels(votes, nclasses_id, threshold=0.5, filter_classes=None, out=None):
        return torch.as_tensor(orc.segment(votes.numpy(), nclasses_id, threshold, filter_classes))

    monkeypatch.setattr(voting, "require_cuda", lambda: torch.device("cpu"))
    monkeypatch.setattr(eng, "resize_nearest", resize_nearest)
    monkeypatch.setattr(eng, "vote_uv2pt", vote_uv2pt)
    monkeypatch.setattr(eng, "vote_finalize", lambda packed: packed)
    monkeypatch.setattr(eng, "resolve_labels", resolve_labels)
    return voting


def _write_scan(root, g):
    import cv2
    H, W, N = int(g["H"]), int(g["W"]), int(g["npts"])
    pts = np.random.default_rng(1).random((N, 3))
    (root / "fusion" / "uv2pt").mkdir(parents=True)
    with open(root / "fusion" / "fusion_data.pkl", "wb") as fp:
        pickle.dump({"points": pts, "normals": np.zeros_like(pts), "colors": np.full_like(pts, 0.5), "nmerges": None,
                     "occurences": None, "nframes": len(g["uv2pt"]), "depth_hw": (H, W)}, fp)
    (root / "masks").mkdir()
    for f in range(len(g["uv2pt"])):
        np.save(root / "fusion" / "uv2pt" / f"{f + 1}.npy", g["uv2pt"][f])
        cv2.imwrite(str(root / "masks" / f"{f + 1}.png"), g["masks_big"][f])
    return H, W, N


def test_voting_segmentation_class_host_logic(level_v, tmp_path):
    g = load_golden("g3_levelv")
    H, W, N = _write_scan(tmp_path, g)
    vs = level_v.VotingSegmentation(N, (H, W), tmp_path / "masks", tmp_path / "fusion" / "uv2pt", 133)
    assert vs.nframes == len(g["uv2pt"]) and vs.nclasses == 133
    assert [p.stem for p in vs.mask_files] == [p.stem for p in vs.uv2pt_files]
    votes = vs.vote(resize=True, verbose=False, filename=tmp_path / "v" / "votes.npy")
    assert votes.dtype == np.float64 and np.array_equal(votes, g["votes"].astype(np.float64))
    assert np.array_equal(np.load(tmp_path / "v" / "votes.npy"), votes)
    assert np.array_equal(vs.segment(), orc.segment(g["votes"], 133, 0.5, None))
    assert np.array_equal(vs.segment(0.3, [1, 0, 5]), g["seg_alias"])          # reference-produced (aliasing filter order)
    vs.zero()
    assert vs.votes.sum() == 0
    # reload from file: nclasses becomes the column count (voting.py:40), which changes the "unclassified" id
    vs2 = level_v.VotingSegmentation(None, None, None, None, None, votes_file=tmp_path / "v" / "votes.npy")
    assert vs2.nclasses == 134
    assert np.array_equal(vs2.segment(0.75), orc.segment(g["votes"], 134, 0.75, None))
    vs2.votes = g["votes"].astype(np.float64) * 2                      # setter keeps the reference's attribute assignable
    assert np.array_equal(vs2.segment(0.75), orc.segment(g["votes"], 134, 0.75, None))
    with pytest.raises(Exception):
        vs2.votes = g["votes"].astype(np.float64) + 0.5                # fractional votes have no GPU representation


def test_get3dseg_segment_and_remove_classes_host_logic(level_v, tmp_path):
    g3 = importlib.import_module(PKG_NAME + ".get3DSeg")
    g = load_golden("g3_levelv")
    _write_scan(tmp_path, g)
    votes, classes = g3.segment(tmp_path, tmp_path / "masks", verbose=False)       # reference defaults
    assert np.array_equal(votes, g["votes"].astype(np.float64)) and np.array_equal(classes, g["seg_default"])
    assert np.array_equal(np.load(tmp_path / "segmentation" / "votes.npy"), votes)
    assert np.array_equal(np.load(tmp_path / "segmentation" / "classes.npy"), classes)
    assert (tmp_path / "segmentation" / "info.json").is_file() and (tmp_path / "segmentation" / "pcd.ply").is_file()
    _, classes2 = g3.segment(tmp_path, tmp_path / "masks", threshold=0.3, filter_classes=[1, 0, 5], verbose=False)
    assert np.array_equal(classes2, g["seg_alias"])
    keep = [0, 1, 2, 3, 4]
    mask = g3.remove_classes(tmp_path, tmp_path / "masks", keep, verbose=False)
    cls = orc.segment(g["votes"], 134, 0.75, None)
    removed = np.append(np.setdiff1d(np.arange(133), keep), [133, 134])
    assert mask.dtype == bool and np.array_equal(mask, ~np.isin(cls, removed))
    assert np.array_equal(np.load(tmp_path / "segmentation" / "remaining_mask.npy"), mask)
