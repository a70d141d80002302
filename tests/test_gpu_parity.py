"""GPU parity tests: the CUDA path (called through the C ABI) against the golden vectors produced by the unmodified
reference and against the numpy oracle on seeded scenes.  Integer outcomes must be bit-exact."""
import importlib

import numpy as np
import pytest
import torch

from conftest import PKG_NAME, load_golden, small_scene
from oracle import f3d_oracle as orc

pytestmark = pytest.mark.gpu


def dev(a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


# ---------------------------------------------------------------------------------------------------------------------
# frame table, projection, cull (a-1, a-3, a-4)
# ---------------------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("tag", ["640", "1920", "3840"])
def test_frame_table_matches_oracle_bitwise(engine, tag):
    g = load_golden(f"g1_{tag}")
    W, H = int(g["W"]), int(g["H"])
    tab = engine.FrameTable(g["K"], W, H, g["wxyz"], g["t"], 4.0)
    eyes, look, nrm = [x.cpu().numpy() for x in tab.export()]
    oe, ol, on = orc.frustum_data(g["K"], W, H, g["wxyz"], g["t"])
    assert np.array_equal(eyes, oe) and np.array_equal(look, ol) and np.array_equal(nrm, on)
    # and the reference's own numbers up to its BLAS rounding
    np.testing.assert_allclose(nrm, g["face_normals"], rtol=0, atol=4e-16)


@pytest.mark.parametrize("tag", ["640", "1920", "3840"])
def test_points2pixel_and_cull_golden(engine, tag):
    cam = importlib.import_module(PKG_NAME + ".Fusion3DSeg.camera_utils")
    isec = importlib.import_module(PKG_NAME + ".Fusion3DSeg.intersections")
    g = load_golden(f"g1_{tag}")
    p = g["points"].astype(np.float64)
    K, W, H = g["K"], int(g["W"]), int(g["H"])
    for j in range(len(g["t"])):
        uv = cam.points2pixel(p, K, g["wxyz"][j], g["t"][j])
        assert uv.dtype == np.int32 and uv.shape == (2, len(p))
        with np.errstate(all="ignore"):
            ouv = orc.points2pixel(p, K, g["wxyz"][j], g["t"][j])
        assert np.array_equal(uv, ouv)                              # everywhere, including behind the camera
        _, _, h2 = orc.project_homogeneous(p, K, g["wxyz"][j], g["t"][j])
        front = h2 > 1e-3
        assert np.array_equal(uv[:, front], g["uv"][j][:, front])   # the reference's output
        pp, pn = orc.frame_planes(g["eyes"][j], g["lookats"][j], g["face_normals"][j], 4.0)
        inside = isec.point_inside_polyhedra(p, pp, pn)
        assert inside.dtype == bool and np.array_equal(inside, g["inside"][j])


# ---------------------------------------------------------------------------------------------------------------------
# kernel (1) fused project + z-test + gather + vote, uv2pt writer
# ---------------------------------------------------------------------------------------------------------------------

def run_fused(engine, s, depth=None, radius=0.05, zmin=0.1, zmax=None, audit=False, chunks=None, nclasses1=134):
    zmax = s["zmax"] if zmax is None else zmax
    tab = engine.FrameTable(s["K"], s["W"], s["H"], s["wxyz"], s["t"], zmax)
    p4 = engine.pack_points(s["points"])
    d = dev(s["depths"] if depth is None else depth)
    m = dev(s["masks"])
    stats = engine.new_stats()
    F = len(s["t"])
    if chunks is None:
        votes = engine.fuse_project_vote(p4, tab, d, m, nclasses1, radius, zmin, zmax, stats=stats, audit=audit)
    else:
        votes = None
        for (a, b) in chunks:
            votes = engine.fuse_project_vote(p4, tab, d[a:b], m[a:b], nclasses1, radius, zmin, zmax, votes=votes,
                                             accumulate=votes is not None, stats=stats, audit=audit, frame_begin=a,
                                             frame_end=b)
    torch.cuda.synchronize()
    return votes.cpu().numpy(), engine.stats_dict(stats), tab, p4


def test_level_p_golden(engine):
    g = load_golden("g2_levelp")
    s = dict(points=g["points"], K=g["K"], W=int(g["W"]), H=int(g["H"]), wxyz=g["wxyz"], t=g["t"], depths=g["depths"],
             masks=g["masks"], zmax=float(g["zmax"]))
    votes, st, tab, p4 = run_fused(engine, s)
    assert votes.dtype == np.int32 and np.array_equal(votes, g["votes"])
    assert st["seen"] == int(g["votes"].sum()) and st["audit_bad"] == 0
    votes_a, st_a, _, _ = run_fused(engine, s, audit=True)
    assert np.array_equal(votes_a, g["votes"]) and st_a["audit_bad"] == 0
    uv2pt = engine.fuse_uv2pt(p4, tab, dev(g["depths"]), 0.05, 0.1, 4.0).cpu().numpy()
    assert np.array_equal(uv2pt, g["uv2pt"])
    lab = engine.resolve_labels(dev(votes), 133, 0.5, [86, 114, 115]).cpu().numpy()
    assert lab.dtype == np.int64 and np.array_equal(lab, g["seg_default"])
    assert np.array_equal(engine.resolve_labels(dev(votes), 133, 0.5, None).cpu().numpy(), g["seg_all"])


@pytest.mark.parametrize("W,H,N,F,seed", [(160, 120, 30000, 6, 7), (640, 480, 50001, 5, 11), (1920, 1440, 40000, 4, 13),
                                          (3840, 2160, 40000, 3, 17), (64, 48, 255, 3, 19), (64, 48, 1, 2, 23)])
def test_level_p_vs_oracle(engine, scenes, W, H, N, F, seed):
    s = small_scene(scenes, orc, npoints=N, nframes=F, width=W, height=H, seed=seed, border=2, block=16)
    ost = {}
    ov = orc.fuse_project_vote(s["points"], s["K"], W, H, s["wxyz"], s["t"], s["depths"], s["masks"], 134, 0, 0.05, 0.1,
                               s["zmax"], s["zmax"], stats=ost)
    votes, st, _, _ = run_fused(engine, s)
    assert np.array_equal(votes, ov)
    assert st["near_edge"] == ost.get("near_edge_1e-4", 0)          # boundary points are counted exactly
    assert st["seen"] == int(ov.sum())
    votes_a, st_a, _, _ = run_fused(engine, s, audit=True)
    assert np.array_equal(votes_a, ov) and st_a["audit_bad"] == 0   # fp32 never certifies a wrong outcome
    if N > 1000:
        assert ov.sum() > 0 and st["exact"] < 0.2 * max(st["candidates"], 1)


def test_level_p_float32_depth_and_thresholds(engine, scenes):
    s = small_scene(scenes, orc, npoints=30000, nframes=5, width=320, height=240, seed=29, border=0)
    dm = (s["depths"].astype(np.float32) * np.float32(0.001)).astype(np.float32)
    for radius, zmin, zmax in [(0.05, 0.1, 4.0), (0.004, 0.5, 2.5), (0.3, 0.0, 3.0)]:
        ov = orc.fuse_project_vote(s["points"], s["K"], s["W"], s["H"], s["wxyz"], s["t"], dm, s["masks"], 134, 1, radius,
                                   zmin, zmax, 4.0)
        tab = engine.FrameTable(s["K"], s["W"], s["H"], s["wxyz"], s["t"], 4.0)
        votes = engine.fuse_project_vote(engine.pack_points(s["points"]), tab, dev(dm), dev(s["masks"]), 134, radius, zmin,
                                         zmax).cpu().numpy()
        assert np.array_equal(votes, ov)
        ov16 = orc.fuse_project_vote(s["points"], s["K"], s["W"], s["H"], s["wxyz"], s["t"], s["depths"], s["masks"], 134, 0,
                                     radius, zmin, zmax, 4.0)
        v16 = engine.fuse_project_vote(engine.pack_points(s["points"]), tab, dev(s["depths"]), dev(s["masks"]), 134, radius,
                                       zmin, zmax).cpu().numpy()
        assert np.array_equal(v16, ov16)


def test_level_p_chunked_accumulate_and_empty(engine, scenes):
    s = small_scene(scenes, orc, npoints=20000, nframes=7, width=160, height=120, seed=31)
    full, _, tab, p4 = run_fused(engine, s)
    a, _, _, _ = run_fused(engine, s, chunks=[(0, 3), (3, 7)])
    b, _, _, _ = run_fused(engine, s, chunks=[(4, 7), (0, 2), (2, 4)])      # votes commute over frames
    assert np.array_equal(full, a) and np.array_equal(full, b)
    # zero frames: vote tensor is overwritten with zeros (no memset needed by the caller)
    z = torch.full((len(s["points"]), 134), 7, dtype=torch.int32, device="cuda")
    engine.fuse_project_vote(p4, tab, dev(s["depths"][:0]), dev(s["masks"][:0]), 134, votes=z, accumulate=False,
                             frame_begin=0, frame_end=0)
    assert int(z.abs().sum()) == 0
    # all-zero depth (no valid pixel) -> no votes
    v0 = engine.fuse_project_vote(p4, tab, dev(np.zeros_like(s["depths"])), dev(s["masks"]), 134).cpu().numpy()
    assert v0.sum() == 0


def test_fused_resolve_epilogue(engine, scenes):
    """f3d_fuse_project_vote_resolve: labels straight from the on-chip histograms == segment(votes)."""
    s = small_scene(scenes, orc, npoints=30011, nframes=6, width=320, height=240, seed=53, block=16)
    ov = orc.fuse_project_vote(s["points"], s["K"], s["W"], s["H"], s["wxyz"], s["t"], s["depths"], s["masks"], 134, 0, 0.05,
                               0.1, 4.0, 4.0)
    tab = engine.FrameTable(s["K"], s["W"], s["H"], s["wxyz"], s["t"], 4.0)
    p4 = engine.pack_points(s["points"])
    d, m = dev(s["depths"]), dev(s["masks"])
    for thr, fc, ncls in [(0.5, None, 133), (0.5, [86, 114, 115], 133), (0.3, [1, 0, 5], 133), (0.75, None, 134),
                          (0.0, list(range(133, -1, -1)), 133)]:
        votes, labels = engine.fuse_project_vote_resolve(p4, tab, d, m, 134, ncls, 0.05, 0.1, 4.0, thr, fc)
        assert np.array_equal(votes.cpu().numpy(), ov)
        assert np.array_equal(labels.cpu().numpy(), orc.segment(ov, ncls, thr, fc)), (thr, fc, ncls)
    none_votes, labels = engine.fuse_project_vote_resolve(p4, tab, d, m, 134, 133, 0.05, 0.1, 4.0, 0.5, None, want_votes=False)
    assert none_votes is None and np.array_equal(labels.cpu().numpy(), orc.segment(ov, 133, 0.5, None))


def test_packed_uint16_votes(engine, scenes):
    """The multi-GPU exchange format: uint16 counters, pairs summed as int32, resolved without widening."""
    s = small_scene(scenes, orc, npoints=20011, nframes=6, width=320, height=240, seed=59, block=16)
    tab = engine.FrameTable(s["K"], s["W"], s["H"], s["wxyz"], s["t"], 4.0)
    p4 = engine.pack_points(s["points"])
    d, m = dev(s["depths"]), dev(s["masks"])
    v32 = engine.fuse_project_vote(p4, tab, d, m, 134, 0.05, 0.1, 4.0)
    v16 = engine.fuse_project_vote(p4, tab, d, m, 134, 0.05, 0.1, 4.0, packed_u16=True)
    assert v16.dtype == torch.uint16 and torch.equal(v16.to(torch.int32), v32)
    a = engine.fuse_project_vote(p4, tab, d[:3], m[:3], 134, 0.05, 0.1, 4.0, frame_begin=0, frame_end=3, packed_u16=True)
    b = engine.fuse_project_vote(p4, tab, d[3:], m[3:], 134, 0.05, 0.1, 4.0, frame_begin=3, frame_end=6, packed_u16=True)
    summed = (a.view(torch.int32) + b.view(torch.int32)).view(torch.uint16)     # what the int32 reduce-scatter computes
    assert torch.equal(summed.to(torch.int32), v32)
    acc = engine.fuse_project_vote(p4, tab, d[3:], m[3:], 134, 0.05, 0.1, 4.0, votes=a.clone(), accumulate=True, frame_begin=3,
                                   frame_end=6)
    assert torch.equal(acc.to(torch.int32), v32)
    for thr, fc in [(0.5, None), (0.5, [86, 114, 115]), (0.3, [1, 0, 5])]:
        assert torch.equal(engine.resolve_labels(v16, 133, thr, fc), engine.resolve_labels(v32, 133, thr, fc))


def test_unsorted_cloud_same_votes(engine, scenes):
    s = small_scene(scenes, orc, npoints=20000, nframes=4, width=160, height=120, seed=37)
    base, _, _, _ = run_fused(engine, s)
    perm = np.random.default_rng(0).permutation(len(s["points"]))
    s2 = dict(s, points=np.ascontiguousarray(s["points"][perm]))
    shuf, _, _, _ = run_fused(engine, s2)
    assert np.array_equal(shuf, base[perm])


# ---------------------------------------------------------------------------------------------------------------------
# kernel (2) z-buffer splat
# ---------------------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("W,H,N,F,seed", [(160, 120, 30000, 5, 41), (1920, 1440, 30000, 2, 43), (96, 64, 60000, 3, 47)])
def test_zbuffer_splat_vs_oracle(engine, scenes, W, H, N, F, seed):
    s = small_scene(scenes, orc, npoints=N, nframes=F, width=W, height=H, seed=seed, border=3)
    tab = engine.FrameTable(s["K"], W, H, s["wxyz"], s["t"], s["zmax"])
    st = engine.new_stats()
    d = engine.zbuffer_splat(engine.pack_points(s["points"]), tab, border=3, stats=st).cpu().numpy()
    assert d.dtype == np.uint16 and np.array_equal(d, s["depths"])
    d_a = engine.zbuffer_splat(engine.pack_points(s["points"]), tab, border=3, stats=st, audit=True).cpu().numpy()
    assert np.array_equal(d_a, s["depths"]) and engine.stats_dict(st)["audit_bad"] == 0


# ---------------------------------------------------------------------------------------------------------------------
# level V vote + resize, kernel (3) resolve
# ---------------------------------------------------------------------------------------------------------------------

def test_level_v_golden_through_reference_class(engine, tmp_path):
    import cv2
    voting = importlib.import_module(PKG_NAME + ".Fusion3DSeg.segUtils.voting")
    g = load_golden("g3_levelv")
    H, W = int(g["H"]), int(g["W"])
    (tmp_path / "masks").mkdir()
    (tmp_path / "uv2pt").mkdir()
    for f in range(len(g["uv2pt"])):
        np.save(tmp_path / "uv2pt" / f"{f + 1}.npy", g["uv2pt"][f])
        cv2.imwrite(str(tmp_path / "masks" / f"{f + 1}.png"), g["masks_big"][f])
    voter = voting.VotingSegmentation(int(g["npts"]), (H, W), tmp_path / "masks", tmp_path / "uv2pt", 133)
    assert voter.nframes == len(g["uv2pt"]) and voter.votes.shape == (int(g["npts"]), 134)
    v = voter.vote(resize=True, filename=tmp_path / "seg" / "votes.npy")
    assert v.dtype == np.float64 and np.array_equal(v, g["votes"].astype(np.float64))
    assert np.array_equal(np.load(tmp_path / "seg" / "votes.npy"), v)
    cases = {"seg_default": (0.5, [86, 114, 115]), "seg_all": (0.5, None), "seg_t075": (0.75, None),
             "seg_alias": (0.3, [1, 0, 5]), "seg_t0": (0.0, [3, 2, 1, 0])}
    for k, (thr, fc) in cases.items():
        out = voter.segment(thr, fc)
        assert out.dtype == np.int64 and np.array_equal(out, g[k]), k
    # votes= argument, votes_file constructor (nclasses quirk: 134 when loaded, voting.py:40)
    assert np.array_equal(voter.segment(0.5, None, votes=v), g["seg_all"])
    v2 = voting.VotingSegmentation(None, None, None, None, None, votes_file=tmp_path / "seg" / "votes.npy")
    assert v2.nclasses == 134
    ref_quirk = orc.segment(g["votes"], 134, 0.75, None)
    assert np.array_equal(v2.segment(0.75, None), ref_quirk)
    # second vote() call accumulates like the reference (votes double)
    voter.vote(resize=True)
    assert np.array_equal(voter.votes, 2.0 * g["votes"])
    voter.zero()
    assert voter.votes.sum() == 0


def test_resize_nearest_golden(engine):
    g = load_golden("g4_resize")
    for k in range(4):
        d = g[f"dst{k}"]
        out = engine.resize_nearest(dev(g[f"src{k}"][None]), d.shape[0], d.shape[1]).cpu().numpy()[0]
        assert np.array_equal(out, d)


def test_level_v_dedup_random(engine):
    rng = np.random.default_rng(5)
    N, C1, F, npix = 5000, 134, 9, 40000
    uv = rng.integers(-1, 600, (F, npix)).astype(np.int32)            # heavy duplication: 600 points, 40k pixels
    uv[rng.random((F, npix)) < 0.3] = -1
    mask = rng.integers(0, 6, (F, npix)).astype(np.uint8)
    ov = np.zeros((N, C1), np.int64)
    for f in range(F):
        orc.vote_uv2pt(ov, uv[f], mask[f])
    packed = torch.zeros((N, C1), dtype=torch.int32, device="cuda")
    engine.vote_uv2pt(packed, dev(uv), dev(mask), 1)
    engine.vote_finalize(packed)
    assert np.array_equal(packed.cpu().numpy(), ov)
    assert ov.max() == F


def test_resolve_random_vs_oracle(engine):
    rng = np.random.default_rng(9)
    N, C1 = 20011, 134
    votes = np.zeros((N, C1), np.int32)
    nz = rng.random((N, C1)) < 0.03
    votes[nz] = rng.integers(1, 6, nz.sum())
    votes[:500] = 0                                                   # unvoted points
    votes[500:700, 133] = 9                                           # mostly "unclassified" votes
    votes[700:900, [3, 7]] = 4                                        # exact ties -> first maximum
    dv = dev(votes)
    for thr in (0.0, 0.5, 0.75, 1.0):
        for fc in (None, [86, 114, 115], [1, 0], [7, 3], [5, 5, 2], list(range(133, -1, -1)), [133]):
            for ncls in (133, 134, 2):
                ref = orc.segment(votes, ncls, thr, fc)
                out = engine.resolve_labels(dv, ncls, thr, fc).cpu().numpy()
                assert np.array_equal(out, ref), (thr, fc, ncls)
    # a narrow vote matrix and a single row
    small = rng.integers(0, 3, (7, 5)).astype(np.int32)
    assert np.array_equal(engine.resolve_labels(dev(small), 4, 0.4, None).cpu().numpy(), orc.segment(small, 4, 0.4, None))


# ---------------------------------------------------------------------------------------------------------------------
# kernel (4) box pairs + union-find
# ---------------------------------------------------------------------------------------------------------------------

def test_box_merge_vs_oracle(engine, scenes):
    lo, hi, group, _ = scenes.make_boxes(4000, seed=3, extent=(40.0, 40.0, 6.0), ngroups=4)
    lo[10], hi[10] = lo[11].copy(), lo[11].copy() + 0.0               # degenerate box touching a corner (closed test)
    group[10] = group[11]
    oe = orc.box_pairs_aabb(lo, hi, group)
    assert len(oe) > 500
    e = engine.box_pairs_aabb(lo, hi, group, cap=16)                   # forces the grow-and-retry path
    es = np.unique(e.cpu().numpy().astype(np.int64), axis=0)
    assert len(es) == len(e) and np.array_equal(es, oe)
    lab = engine.union_find(len(lo), e).cpu().numpy()
    assert np.array_equal(lab, orc.union_find_labels(len(lo), oe))
    assert np.array_equal(engine.union_find(5, None).cpu().numpy(), np.arange(5))


def test_obb_contains_vs_oracle(engine):
    rng = np.random.default_rng(2)
    pts = rng.normal(0, 1.0, (20000, 3))
    boxes = []
    for _ in range(5):
        A = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        boxes.append(np.concatenate([rng.normal(0, 0.3, 3), A.reshape(-1), rng.uniform(0.5, 2.0, 3)]))
    boxes = np.stack(boxes)
    out = engine.obb_contains(pts, boxes).cpu().numpy().astype(bool)
    for b in range(5):
        ref = orc.obb_contains(boxes[b, :3], boxes[b, 3:12].reshape(3, 3), boxes[b, 12:], pts)
        assert np.array_equal(out[b], ref)


# ---------------------------------------------------------------------------------------------------------------------
# byte-histogram flush rule: more candidate frames per tile than a uint8 counter can hold
# ---------------------------------------------------------------------------------------------------------------------

def test_byte_histogram_flush_many_views(engine, scenes, monkeypatch):
    """600 frames that all see the same points (3 poses repeated 300 times): single cells reach counts of several
    hundred, so the uint8 shared-memory histogram must flush mid-sweep (first flush overwrites, later ones add) and the
    fused labels must be re-derived from the complete rows.  Checked against 300 x the oracle's 3-frame votes, and
    against the uint16-histogram build of the same kernel."""
    rep = 300
    s = small_scene(scenes, orc, npoints=6001, nframes=3, width=96, height=64, seed=61, block=16)
    ov3 = orc.fuse_project_vote(s["points"], s["K"], s["W"], s["H"], s["wxyz"], s["t"], s["depths"], s["masks"], 134, 0, 0.05,
                                0.1, 4.0, 4.0)
    ov = ov3 * rep
    assert ov.max() > 255
    big = dict(s, wxyz=np.tile(s["wxyz"], (rep, 1)), t=np.tile(s["t"], (rep, 1)), depths=np.tile(s["depths"], (rep, 1, 1)),
               masks=np.tile(s["masks"], (rep, 1, 1)))
    tab = engine.FrameTable(big["K"], big["W"], big["H"], big["wxyz"], big["t"], 4.0)
    p4 = engine.pack_points(big["points"])
    d, m = dev(big["depths"]), dev(big["masks"])
    for hist16 in (False, True):
        if hist16:
            monkeypatch.setenv("F3D_HIST16", "1")
        else:
            monkeypatch.delenv("F3D_HIST16", raising=False)
        st = engine.new_stats()
        votes = engine.fuse_project_vote(p4, tab, d, m, 134, 0.05, 0.1, 4.0, stats=st)
        assert np.array_equal(votes.cpu().numpy(), ov) and engine.stats_dict(st)["seen"] == int(ov.sum())
        v16 = engine.fuse_project_vote(p4, tab, d, m, 134, 0.05, 0.1, 4.0, packed_u16=True)
        assert np.array_equal(v16.to(torch.int32).cpu().numpy(), ov)
        # accumulate on top of an earlier launch (every flush adds)
        acc = engine.fuse_project_vote(p4, tab, d, m, 134, 0.05, 0.1, 4.0, votes=votes.clone(), accumulate=True)
        assert np.array_equal(acc.cpu().numpy(), 2 * ov)
        for thr, fc in [(0.5, None), (0.3, [1, 0, 5])]:
            v, lab = engine.fuse_project_vote_resolve(p4, tab, d, m, 134, 133, 0.05, 0.1, 4.0, thr, fc)
            assert np.array_equal(v.cpu().numpy(), ov)
            assert np.array_equal(lab.cpu().numpy(), orc.segment(ov, 133, thr, fc))
        # labels without a vote tensor: nowhere to flush to, so the library runs the uint16 histogram
        none_votes, lab = engine.fuse_project_vote_resolve(p4, tab, d, m, 134, 133, 0.05, 0.1, 4.0, 0.5, None, want_votes=False)
        assert none_votes is None and np.array_equal(lab.cpu().numpy(), orc.segment(ov, 133, 0.5, None))
        # audit mode evaluates every candidate in fp64 inside the sweep
        va = engine.fuse_project_vote(p4, tab, d, m, 134, 0.05, 0.1, 4.0, stats=st, audit=True)
        assert np.array_equal(va.cpu().numpy(), ov) and engine.stats_dict(st)["audit_bad"] == 0
    monkeypatch.delenv("F3D_HIST16", raising=False)


def test_uint16_histogram_build_matches(engine, scenes, monkeypatch):
    """The uint16-histogram instantiation (used for labels-only launches over many frames) stays bit-identical."""
    s = small_scene(scenes, orc, npoints=20011, nframes=6, width=320, height=240, seed=67, block=16)
    base, _, _, _ = run_fused(engine, s)
    monkeypatch.setenv("F3D_HIST16", "1")
    alt, _, _, _ = run_fused(engine, s)
    monkeypatch.delenv("F3D_HIST16", raising=False)
    assert np.array_equal(base, alt)


# ---------------------------------------------------------------------------------------------------------------------
# multi-GPU vote exchange, emulated on one device: every "rank" fuses its frame shard into the owners' receive buffers
# ---------------------------------------------------------------------------------------------------------------------

def _emulated_exchange(engine, s, world, nclasses_id=133, thr=0.5, fc=None, sub_rows=None, sub_cap=None, compact=False):
    parallel = importlib.import_module(PKG_NAME + ".parallel")
    N, C1, F = len(s["points"]), 134, len(s["t"])
    tab = engine.FrameTable(s["K"], s["W"], s["H"], s["wxyz"], s["t"], s["zmax"])
    p4 = engine.pack_points(s["points"])
    d, m = dev(s["depths"]), dev(s["masks"])
    nreg, nsub, nfix, nlev = engine.exchange_constants()
    per = parallel.shard_points(N, world)
    blocks = per // 32
    sub_rows = sub_rows or max(64, -(-blocks * 40 // nreg))
    sub_cap = sub_cap or 256
    # owner o: queue [world][nsub][sub_cap] + counts [world][nsub] + directory [world][blocks] + records [world][nreg*sub_rows] x 64 B
    queues = [torch.zeros(world * nsub * sub_cap, dtype=torch.int64, device="cuda") for _ in range(world)]
    counts = [torch.zeros(world * nsub, dtype=torch.int32, device="cuda") for _ in range(world)]
    dirs = [torch.full((world * blocks * nlev,), -1, dtype=torch.int64, device="cuda") for _ in range(world)]  # garbage: must be overwritten
    slots = [torch.full((world * nreg * sub_rows * 32,), 0x7b7b, dtype=torch.uint16, device="cuda") for _ in range(world)]
    overflow = torch.zeros(1, dtype=torch.int32, device="cuda")
    st = engine.new_stats()
    nrec = 0
    for r in range(world):                                              # source ranks, one after the other
        fb, fe = parallel.frame_shard(F, r, world)
        cursors = torch.zeros(world * (nreg + nsub), dtype=torch.int32, device="cuda")
        qptrs = np.array([queues[o].data_ptr() + r * nsub * sub_cap * 8 for o in range(world)], dtype=np.uint64)
        sptrs = np.array([slots[o].data_ptr() + r * nreg * sub_rows * 64 for o in range(world)], dtype=np.uint64)
        dptrs = np.array([dirs[o].data_ptr() + r * blocks * nlev * 8 for o in range(world)], dtype=np.uint64)
        engine.fuse_project_vote_exchange(p4, tab, d[fb:fe], m[fb:fe], C1, world, per, sptrs, dptrs, qptrs, sub_rows, sub_cap, cursors,
                                          overflow, 0.05, 0.1, s["zmax"], stats=st, frame_begin=fb, frame_end=fe)
        cptrs = np.array([counts[o].data_ptr() for o in range(world)], dtype=np.uint64)
        engine.exchange_publish(cursors, cptrs, r, sub_cap)
        nrec += int(cursors[:world * nreg].sum())
    torch.cuda.synchronize()
    assert int(overflow.item()) == 0
    votes, labels = [], []
    # the label all-gather rides inside the owner kernels: every "rank" keeps an int16 copy of all labels
    full16 = [torch.full((per * world,), -7, dtype=torch.int16, device="cuda") for _ in range(world)]
    lptrs = np.array([t.data_ptr() for t in full16], dtype=np.uint64)
    for o in range(world):                                              # owner ranks
        rows = max(0, min(per, N - o * per))
        shard = torch.full((per, C1), 77, dtype=torch.int32, device="cuda")       # garbage: every cell must be written
        lab = torch.zeros(per, dtype=torch.int64, device="cuda")
        if rows:
            engine.exchange_merge(slots[o], dirs[o], world, sub_rows, per, rows, C1, nclasses_id, thr, fc, votes=shard, labels=lab,
                                  peer_labels16=lptrs, first_point=o * per)
            engine.exchange_queue_apply(queues[o], counts[o], world, sub_cap, shard, rows, nclasses_id, lab, thr, fc,
                                        peer_labels16=lptrs, first_point=o * per)
        votes.append(shard[:rows])
        labels.append(lab[:rows])
    torch.cuda.synchronize()
    for t in full16:                                                    # every copy holds every owner's labels
        assert torch.equal(t[:N].to(torch.int64), torch.cat(labels))
        assert bool((t[N:] == -7).all())
    return torch.cat(votes).cpu().numpy(), torch.cat(labels).cpu().numpy(), engine.stats_dict(st), [int(c.sum()) for c in counts], nrec


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_vote_exchange_emulated_ranks(engine, scenes, world):
    s = small_scene(scenes, orc, npoints=20011, nframes=9, width=320, height=240, seed=71, block=16)
    ov = orc.fuse_project_vote(s["points"], s["K"], s["W"], s["H"], s["wxyz"], s["t"], s["depths"], s["masks"], 134, 0, 0.05,
                               0.1, 4.0, 4.0)
    for thr, fc in [(0.5, None), (0.3, [1, 0, 5])]:
        votes, labels, st, nq, nrec = _emulated_exchange(engine, s, world, thr=thr, fc=fc)
        assert np.array_equal(votes, ov)
        assert np.array_equal(labels, orc.segment(ov, 133, thr, fc))
        assert st["seen"] == int(ov.sum())
        assert sum(nq) < 0.02 * (ov > 0).sum() and nrec > 0             # only deferred fp64 votes use the queue


def test_vote_exchange_spills_and_flushes(engine, scenes):
    """300 frames (3 poses x 100) with a different random mask each: points collect far more than 32 distinct classes
    from one source (long lists: several staging chunks, re-read from the histogram row) and tiles sweep more than 235
    candidate frames (mid-sweep flush: the first flush writes the record, later ones go to the queue).  A second pass
    with tiny record sub-regions forces the region-full fallback to the queue."""
    rep = 100
    base = small_scene(scenes, orc, npoints=5003, nframes=3, width=96, height=64, seed=73, block=16)
    F = 3 * rep
    s = dict(base, wxyz=np.tile(base["wxyz"], (rep, 1)), t=np.tile(base["t"], (rep, 1)), depths=np.tile(base["depths"], (rep, 1, 1)),
             masks=scenes.block_masks((base["H"], base["W"]), F, seed=79, block=16))
    ov = orc.fuse_project_vote(s["points"], s["K"], s["W"], s["H"], s["wxyz"], s["t"], s["depths"], s["masks"], 134, 0, 0.05,
                               0.1, 4.0, 4.0)
    assert ((ov > 0).sum(axis=1) > 32).sum() > 100                      # many points have long lists
    for world, sub_rows in ((1, 256), (2, 256), (2, 2)):
        votes, labels, _, nq, nrec = _emulated_exchange(engine, s, world, sub_rows=sub_rows, sub_cap=1 << 14)
        assert np.array_equal(votes, ov) and sum(nq) > 0
        assert np.array_equal(labels, orc.segment(ov, 133, 0.5, None))


@pytest.mark.parametrize("world", [2, 8])
def test_vote_exchange_compacted_launch(engine, scenes, world):
    """F3D_FUSE_COMPACT: the fused kernel runs only over the super-tiles a rank's frames can see; the directory entries of
    the skipped ones (pre-filled with garbage here) are cleared by dead_directory_kernel.  > 32 frames per rank so that the
    first cull level runs, unseen super-tiles before / between / after the scene, ragged last super-tile."""
    s = dict(small_scene(scenes, orc, npoints=9000, nframes=40 * world, width=96, height=72, seed=13))
    far = np.random.default_rng(0).uniform(-1, 1, (3 * 4096 + 777, 3)).astype(np.float32) + np.float32(500.0)
    s["points"] = np.concatenate([far[:4096], s["points"][:5000], far[4096:8192], s["points"][5000:], far[8192:]])
    ov = orc.fuse_project_vote(s["points"], s["K"], s["W"], s["H"], s["wxyz"], s["t"], s["depths"], s["masks"], 134, 0, 0.05,
                               0.1, 4.0, 4.0)
    assert ov[:4096].sum() == 0 and ov.sum() > 1000
    plain = _emulated_exchange(engine, s, world)
    comp = _emulated_exchange(engine, s, world, compact=True)
    for votes, labels, st, nq, nrec in (plain, comp):
        assert np.array_equal(votes, ov)
        assert np.array_equal(labels, orc.segment(ov, 133, 0.5, None))
        assert st["seen"] == int(ov.sum())
    assert comp[2]["candidates"] == plain[2]["candidates"]
