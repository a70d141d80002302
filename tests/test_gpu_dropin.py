"""GPU: the reference-shaped entry points (get3DSeg.segment / remove_classes, Fusion helpers, merge_bb, merge_boxes)
against the oracle, plus size-independent properties at the full BASELINE sizes (C1 scene, C5 boxes)."""
import importlib
import pickle

import numpy as np
import pytest
import torch

from conftest import PKG_NAME, load_golden, small_scene
from oracle import f3d_oracle as orc

pytestmark = pytest.mark.gpu


def _write_fusion_dir(root, points, nframes, depth_hw, with_adj=False):
    (root / "fusion" / "uv2pt").mkdir(parents=True)
    data = {"points": points, "normals": np.zeros_like(points), "colors": np.full_like(points, 0.5), "nmerges": None,
            "occurences": None, "nframes": nframes, "depth_hw": depth_hw}
    with open(root / "fusion" / "fusion_data.pkl", "wb") as fp:
        pickle.dump(data, fp)
    if with_adj:
        with open(root / "fusion" / "adj.pkl", "wb") as fp:
            pickle.dump(np.array([np.array([0])] * len(points), dtype=object), fp)


def test_get3dseg_segment_and_remove_classes(engine, tmp_path):
    import cv2
    g3 = importlib.import_module(PKG_NAME + ".get3DSeg")
    g = load_golden("g3_levelv")
    H, W, N = int(g["H"]), int(g["W"]), int(g["npts"])
    pts = np.random.default_rng(1).random((N, 3))
    _write_fusion_dir(tmp_path, pts, len(g["uv2pt"]), (H, W))
    (tmp_path / "masks").mkdir()
    for f in range(len(g["uv2pt"])):
        np.save(tmp_path / "fusion" / "uv2pt" / f"{f + 1}.npy", g["uv2pt"][f])
        cv2.imwrite(str(tmp_path / "masks" / f"{f + 1}.png"), g["masks_big"][f])
    votes, classes = g3.segment(tmp_path, tmp_path / "masks", verbose=False)       # reference defaults
    assert np.array_equal(votes, g["votes"].astype(np.float64)) and np.array_equal(classes, g["seg_default"])
    assert np.array_equal(np.load(tmp_path / "segmentation" / "votes.npy"), votes)
    assert np.array_equal(np.load(tmp_path / "segmentation" / "classes.npy"), classes)
    assert (tmp_path / "segmentation" / "info.json").is_file() and (tmp_path / "segmentation" / "pcd.ply").is_file()
    votes2, classes2 = g3.segment(tmp_path, tmp_path / "masks", threshold=0.3, filter_classes=[1, 0, 5], verbose=False)
    assert np.array_equal(classes2, g["seg_alias"])
    # remove_classes re-uses votes.npy => nclasses quirk (134) of voting.py:40, threshold 0.75, no filter
    keep = [0, 1, 2, 3, 4]
    mask = g3.remove_classes(tmp_path, tmp_path / "masks", keep, verbose=False)
    cls = orc.segment(g["votes"], 134, 0.75, None)
    removed = np.append(np.setdiff1d(np.arange(133), keep), [133, 134])
    assert mask.dtype == bool and np.array_equal(mask, ~np.isin(cls, removed))
    assert np.array_equal(np.load(tmp_path / "segmentation" / "remaining_mask.npy"), mask)


def test_fusion_helpers(engine, scenes, tmp_path):
    fusion = importlib.import_module(PKG_NAME + ".Fusion3DSeg.fusion")
    g = load_golden("g2_levelp")
    W, H = int(g["W"]), int(g["H"])
    eyes, look, spokes, nrm = fusion.Fusion._get_frustum_data(g["K"], W, H, g["wxyz"], g["t"])
    oe, ol, on = orc.frustum_data(g["K"], W, H, g["wxyz"], g["t"])
    assert np.array_equal(eyes, oe) and np.array_equal(look, ol) and np.array_equal(nrm, on)
    assert spokes.shape == (len(g["t"]), 4, 3) and np.array_equal(spokes[:, 2], oe)
    votes, classes = fusion.Fusion.label_fixed_cloud(g["points"], g["K"], W, H, g["wxyz"], g["t"], g["depths"], g["masks"],
                                                     point_range=(0.1, 4), radius=0.05, filter_classes=[86, 114, 115])
    assert votes.dtype == np.float64 and np.array_equal(votes, g["votes"]) and np.array_equal(classes, g["seg_default"])
    names = [str(10 + i) for i in range(len(g["t"]))]
    out = fusion.Fusion.write_uv2pt_fixed_cloud(tmp_path, names, g["points"], g["K"], W, H, g["wxyz"], g["t"], g["depths"],
                                                chunk=3)
    for i, n in enumerate(names):
        assert np.array_equal(np.load(out / f"{n}.npy"), g["uv2pt"][i])
    fusion.Fusion.dump_data(tmp_path, g["points"], nframes=len(names), depth_hw=(H, W))
    loaded = fusion.Fusion.load_data(tmp_path)
    assert len(loaded) == 8 and loaded[7] is None and loaded[6] == (H, W) and np.array_equal(loaded[0], g["points"])
    assert np.array_equal(fusion.FrameData.get_valid(np.array([[0, 0, 0.1], [0, 0, 0.1000001], [0, 0, 4.0], [0, 0, 4.1]]),
                                                     0.1, 4), [False, True, True, False])


def test_merge_bb_sequential_matches_oracle_driver(engine, tmp_path):
    mbb = importlib.import_module(PKG_NAME + ".Fusion3DSeg.merge_intersecting_bb")
    rng = np.random.default_rng(4)
    # 9 instances (id 0 = background): chain 1-2-3 overlapping, 5-6 overlapping, a 3-point instance (7), different parents
    centres = np.array([[0, 0, 0], [1, 0, 0], [1.8, 0, 0], [2.6, 0, 0], [9, 9, 0], [0, 5, 0], [0.7, 5, 0], [5, 5, 5], [3.2, 0, 0]])
    parents = [0, 1, 1, 1, 1, 2, 2, 2, 3]
    pts, ids = [], []
    for i, c in enumerate(centres):
        n = 3 if i == 7 else 300
        pts.append(c + rng.uniform(-0.55, 0.55, (n, 3)) * np.array([1.0, 0.6, 0.3]))
        ids.append(np.full(n, i))
    pts, ids = np.concatenate(pts), np.concatenate(ids).astype(np.int64)
    info = [{"id": i, "category_id": 10 + parents[i], "parent_id": parents[i], "area": int((ids == i).sum())} for i in range(9)]

    def hit_fn(id1, id2, cur_ids):
        def inside(k):
            sel = cur_ids == k
            if sel.sum() < 4:
                return None
            b = mbb.fit_obb(torch.as_tensor(pts[sel]).cuda()).cpu().numpy()
            return orc.obb_contains(b[:3], b[3:12].reshape(3, 3), b[12:], pts)
        a = inside(id1)
        if id2 is None:
            return False if a is None else True
        b = inside(id2)
        return None if b is None else bool((a & b).any())

    ref_info, ref_ids = orc.merge_bb_sequential([dict(d) for d in info], ids.copy(), hit_fn)
    my_info, my_ids = [dict(d) for d in info], ids.copy()

    class Pcd:
        points = pts
    mbb.merge_bb(tmp_path, my_info, my_ids, Pcd())
    assert np.array_equal(my_ids, ref_ids)
    assert [(d["id"], d["area"]) for d in my_info] == [(d["id"], d["area"]) for d in ref_info]
    assert len(my_info) < len(info)                                   # something was merged
    assert np.array_equal(np.load(tmp_path / "panoptic_segmentation" / "ids.npy"), ref_ids)
    assert (tmp_path / "panoptic_segmentation" / "final_info.json").is_file()


def test_merge_boxes_c5_full_size(engine, scenes):
    """BASELINE config 5: 200 k boxes.  Exact edge set and component labels against the numpy sweep oracle, plus
    properties: labels are idempotent minima, every edge joins equal labels, merged areas are conserved."""
    mbb = importlib.import_module(PKG_NAME + ".Fusion3DSeg.merge_intersecting_bb")
    lo, hi, group, area = scenes.make_boxes()
    assert len(lo) == 200_000
    labels, edges, merged = mbb.merge_boxes(lo, hi, group, area)
    lab = labels.cpu().numpy()
    e = np.unique(edges.cpu().numpy().astype(np.int64), axis=0)
    assert len(e) == len(edges)
    oe = orc.box_pairs_aabb(lo, hi, group)
    assert np.array_equal(e, oe) and len(oe) > 100_000
    assert np.array_equal(lab, orc.union_find_labels(len(lo), oe))
    assert np.array_equal(lab[lab], lab) and np.all(lab <= np.arange(len(lab)))
    assert np.all(lab[e[:, 0]] == lab[e[:, 1]]) and np.all(group[e[:, 0]] == group[e[:, 1]])
    assert int(merged.sum()) == int(area.sum()) and np.all(merged.cpu().numpy()[lab != np.arange(len(lab))] == 0)


def test_c1_full_size_properties(engine, scenes):
    """BASELINE config 1 at full size (1 M points x 50 frames 640x480): frame-order commutativity, two-way frame split
    equality (the multi-GPU invariant), vote/stat conservation, fused-resolve == standalone resolve, and a sampled
    oracle check on 20 k points x 6 frames."""
    spec = scenes.CONFIGS["C1"]
    K = scenes.scaled_intrinsics(spec.width, spec.height)
    wxyz, t = scenes.make_poses(spec)
    pts = scenes.make_cloud(spec)
    masks = scenes.block_masks(spec)
    tab = engine.FrameTable(K, spec.width, spec.height, wxyz, t, spec.zmax)
    p4 = engine.pack_points(pts)
    depth = engine.zbuffer_splat(p4, tab, border=10)
    m = torch.as_tensor(masks).cuda()
    st = engine.new_stats()
    votes, labels = engine.fuse_project_vote_resolve(p4, tab, depth, m, 134, 133, 0.05, 0.1, spec.zmax, 0.5, None, stats=st)
    sd = engine.stats_dict(st)
    assert int(votes.sum()) == sd["seen"] > 1_000_000 and int(votes.max()) <= spec.nframes and sd["audit_bad"] == 0
    assert torch.equal(labels, engine.resolve_labels(votes, 133, 0.5, None))
    half = spec.nframes // 2
    a = engine.fuse_project_vote(p4, tab, depth[half:], m[half:], 134, 0.05, 0.1, spec.zmax, frame_begin=half,
                                 frame_end=spec.nframes)
    b = engine.fuse_project_vote(p4, tab, depth[:half], m[:half], 134, 0.05, 0.1, spec.zmax, frame_begin=0, frame_end=half)
    assert torch.equal(a + b, votes)
    # sampled oracle: every 50th point, 6 frames, against the same depth / masks
    sub = np.ascontiguousarray(pts[::50])
    fr = [0, 9, 18, 27, 36, 45]
    d_np, m_np = depth.cpu().numpy()[fr], masks[fr]
    ov = orc.fuse_project_vote(sub, K, spec.width, spec.height, wxyz[fr], t[fr], d_np, m_np, 134, 0, 0.05, 0.1, spec.zmax,
                               spec.zmax)
    tab6 = engine.FrameTable(K, spec.width, spec.height, wxyz[fr], t[fr], spec.zmax)
    gv = engine.fuse_project_vote(engine.pack_points(sub), tab6, torch.as_tensor(d_np).cuda(), torch.as_tensor(m_np).cuda(), 134,
                                  0.05, 0.1, spec.zmax)
    assert np.array_equal(gv.cpu().numpy(), ov) and ov.sum() > 5000
