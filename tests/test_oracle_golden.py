"""CPU: the numpy oracle against the vectors the unmodified reference produced (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import f3d_oracle as orc


@pytest.mark.parametrize("tag", ["640", "1920", "3840"])
def test_projection_cull_frustum(tag):
    g = load_golden(f"g1_{tag}")
    p = g["points"].astype(np.float64)
    K, W, H = g["K"], int(g["W"]), int(g["H"])
    eyes, look, fn = orc.frustum_data(K, W, H, g["wxyz"], g["t"])
    assert np.array_equal(eyes, g["eyes"])
    np.testing.assert_allclose(look, g["lookats"], rtol=0, atol=4e-16)
    np.testing.assert_allclose(fn, g["face_normals"], rtol=0, atol=4e-16)
    for j in range(len(g["t"])):
        pp, pn = orc.frame_planes(eyes[j], look[j], fn[j], 4.0)
        assert np.array_equal(orc.point_inside_polyhedra(p, pp, pn), g["inside"][j])
        with np.errstate(all="ignore"):
            uv = orc.points2pixel(p, K, g["wxyz"][j], g["t"][j])
        _, _, h2 = orc.project_homogeneous(p, K, g["wxyz"][j], g["t"][j])
        front = h2 > 1e-3
        assert front.sum() > 1000
        assert np.array_equal(uv[:, front], g["uv"][j][:, front])
        rot = orc.quat_rotate(orc.quat_inverse(g["wxyz"][j]), p - g["t"][j])[:256]
        np.testing.assert_allclose(rot, g["rot"][j], rtol=0, atol=1e-14)


def test_level_p_votes_uv2pt_segment():
    g = load_golden("g2_levelp")
    W, H = int(g["W"]), int(g["H"])
    v = orc.fuse_project_vote(g["points"], g["K"], W, H, g["wxyz"], g["t"], g["depths"], g["masks"], 134, 0,
                              float(g["radius"]), float(g["zmin"]), float(g["zmax"]), float(g["max_depth"]))
    assert np.array_equal(v, g["votes"])
    assert v.sum() > 1000
    assert np.array_equal(orc.segment(v, 133, 0.5, [86, 114, 115]), g["seg_default"])
    assert np.array_equal(orc.segment(v, 133, 0.5, None), g["seg_all"])
    eyes, look, fn = orc.frustum_data(g["K"], W, H, g["wxyz"], g["t"])
    for f in range(len(g["t"])):
        u = orc.frame_uv2pt(g["points"].astype(np.float64), g["K"], W, H, g["wxyz"][f], g["t"][f], eyes[f], look[f],
                            fn[f], g["depths"][f], 0, 0.05, 0.1, 4.0, 4.0)
        assert np.array_equal(u, g["uv2pt"][f])


def test_level_v_votes_segment_resize():
    g = load_golden("g3_levelv")
    W, H = int(g["W"]), int(g["H"])
    votes = np.zeros((int(g["npts"]), 134), np.int64)
    for f in range(len(g["uv2pt"])):
        m = orc.resize_nearest(g["masks_big"][f], W, H)
        assert np.array_equal(m, g["masks_resized"][f])
        orc.vote_uv2pt(votes, g["uv2pt"][f], m)
    assert np.array_equal(votes, g["votes"])
    cases = {"seg_default": (0.5, [86, 114, 115]), "seg_all": (0.5, None), "seg_t075": (0.75, None),
             "seg_alias": (0.3, [1, 0, 5]), "seg_t0": (0.0, [3, 2, 1, 0])}
    for k, (thr, fc) in cases.items():
        assert np.array_equal(orc.segment(votes, 133, thr, fc), g[k]), k


def test_resize_rule():
    g = load_golden("g4_resize")
    for k in range(4):
        d = g[f"dst{k}"]
        assert np.array_equal(orc.resize_nearest(g[f"src{k}"], d.shape[1], d.shape[0]), d)


def test_segment_aliasing_and_ties():
    votes = np.array([[0, 0, 0, 0], [2, 2, 0, 0], [1, 0, 3, 0], [0, 0, 0, 5], [1, 1, 1, 1]], dtype=np.int64)
    # filter [1, 0]: index 0 -> class 1 -> (i=1) class 0 : everything collapses to 0 (SURVEY a-11)
    out = orc.segment(votes, 3, 0.25, [1, 0])
    assert out.tolist() == [3, 0, 0, 3, 0]
    assert orc.segment(votes, 3, 0.5, None).tolist() == [3, 0, 2, 3, 3]


def test_box_oracle_union_find_chain():
    lo = np.array([[0, 0, 0], [0.9, 0, 0], [1.8, 0, 0], [10, 10, 10], [1.0, 0.5, 0.5]], float)
    hi = lo + 1.0
    group = np.array([0, 0, 0, 0, 1])
    e = orc.box_pairs_aabb(lo, hi, group)
    assert e.tolist() == [[0, 1], [1, 2]]
    assert orc.union_find_labels(5, e).tolist() == [0, 0, 0, 3, 4]
    # closed intervals: touching boxes overlap
    e2 = orc.box_pairs_aabb(np.array([[0, 0, 0], [1, 0, 0]], float), np.array([[1, 1, 1], [2, 1, 1]], float), [0, 0])
    assert e2.tolist() == [[0, 1]]


def test_radius_adjacency_matches_kdtree(scenes):
    """The adjacency restatement (groundwork for the GPU neighbour search) against the very call the
    reference makes: `KDTree(points).query_radius(points, r=2*ds_radius)` (`fusion.py:374-375`), on a room cloud and on a
    lattice whose neighbour distances tie with r exactly (closed ball)."""
    from sklearn.neighbors import KDTree
    spec = scenes.scaled_spec("C1", npoints=6000, nframes=1, seed=17)
    cloud = scenes.make_cloud(spec).astype(np.float64)
    g = np.arange(6) * 0.5
    lattice = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    for pts, r in ((cloud, 0.18), (cloud[:500], 0.05), (lattice, 0.5), (lattice, 0.5 * np.sqrt(2.0)), (lattice[:1], 1.0)):
        indptr, indices = orc.radius_adjacency(pts, r)
        ref = KDTree(pts).query_radius(pts, r=r)
        assert indptr[-1] == sum(len(a) for a in ref)
        for i in range(len(pts)):
            assert np.array_equal(indices[indptr[i]:indptr[i + 1]], np.sort(ref[i]))
    # the instance split consumes it unchanged
    indptr, indices = orc.radius_adjacency(cloud, 0.18)
    classes = (np.floor(cloud[:, 0] / 1.3).astype(np.int64) % 3) * 40 + 5
    ids = orc.split_into_instances(classes, indptr, indices, 133, None, 1)[1]
    assert ids.shape == (len(cloud),)
