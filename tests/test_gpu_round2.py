"""GPU parity tests of the round-2 additions, all through the C ABI: packed frame formats (f3d_pack_frames + fused kernel),
the geometry kernels (radius adjacency, batched box fit, sweep broad phase), many-frames flush paths on 128-point tiles,
and the vote exchange's overflow detection."""
import importlib
import os

import numpy as np
import pytest
import torch

from conftest import PKG_NAME, small_scene
from oracle import f3d_oracle as orc

pytestmark = pytest.mark.gpu


def dev(a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


@pytest.mark.parametrize("fmt", [2, 3])
@pytest.mark.parametrize("wh", [(160, 120), (150, 100)])   # 150 x 100: partial 16 x 16 tiles
def test_pack_frames_layout_and_fused_resize(engine, scenes, fmt, wh):
    W, H = wh
    rng = np.random.default_rng(3)
    F = 3
    depth = rng.integers(0, 65536, (F, H, W)).astype(np.uint16)
    mask2 = rng.integers(0, 134, (F, 2 * H + 1, 2 * W - 3)).astype(np.uint8)        # another resolution: resized on the fly
    pk = engine.pack_frames(dev(depth), dev(mask2), fmt)
    tex = pk.texels.cpu().numpy().view(np.uint32)
    mask = np.stack([orc.resize_nearest(mask2[f], W, H) for f in range(F)])         # oracle of cv2.resize(INTER_NEAREST), voting.py:93
    want = depth.astype(np.uint32) | (mask.astype(np.uint32) << 16)
    if fmt == 2:
        assert tex.shape == (F, H * W) and np.array_equal(tex.reshape(F, H, W), want)
    else:
        tx, ty = -(-W // 16), -(-H // 16)
        assert tex.shape == (F, tx * ty * 256)
        t4 = tex.reshape(F, ty, tx, 16, 16).transpose(0, 1, 3, 2, 4).reshape(F, ty * 16, tx * 16)
        assert np.array_equal(t4[:, :H, :W], want)
        assert not t4[:, H:, :].any() and not t4[:, :, W:].any()                    # padding texels are zero


@pytest.mark.parametrize("fmt", [2, 3])
@pytest.mark.parametrize("wh", [(160, 120), (200, 152), (150, 100)])
def test_fused_packed_formats_match_oracle(engine, scenes, fmt, wh):
    s = small_scene(scenes, orc, npoints=30000, nframes=7, width=wh[0], height=wh[1], seed=11)
    ov = orc.fuse_project_vote(s["points"], s["K"], s["W"], s["H"], s["wxyz"], s["t"], s["depths"], s["masks"], 134, 0, 0.05, 0.1, 4.0, 4.0)
    tab = engine.FrameTable(s["K"], s["W"], s["H"], s["wxyz"], s["t"], 4.0)
    p4 = engine.pack_points(s["points"])
    pk = engine.pack_frames(dev(s["depths"]), dev(s["masks"]), fmt)
    for audit in (False, True):
        st = engine.new_stats()
        votes, labels = engine.fuse_project_vote_resolve(p4, tab, pk, None, 134, 133, 0.05, 0.1, 4.0, 0.5, None, stats=st, audit=audit)
        assert np.array_equal(votes.cpu().numpy(), ov)
        assert np.array_equal(labels.cpu().numpy(), orc.segment(ov, 133, 0.5, None))
        sd = engine.stats_dict(st)
        assert sd["seen"] == ov.sum() and sd["audit_bad"] == 0
    # labels only (no vote tensor), chunked accumulation, and the uv2pt writer read the same texels
    _, l2 = engine.fuse_project_vote_resolve(p4, tab, pk, None, 134, 133, 0.05, 0.1, 4.0, 0.3, [5, 86, 2], want_votes=False)
    assert np.array_equal(l2.cpu().numpy(), orc.segment(ov, 133, 0.3, [5, 86, 2]))
    acc = engine.fuse_project_vote(p4, tab, pk.slice(0, 3), None, 134, 0.05, 0.1, 4.0, frame_begin=0, frame_end=3)
    acc = engine.fuse_project_vote(p4, tab, pk.slice(3, 7), None, 134, 0.05, 0.1, 4.0, votes=acc, accumulate=True, frame_begin=3, frame_end=7)
    assert np.array_equal(acc.cpu().numpy(), ov)
    uv = engine.fuse_uv2pt(p4, tab, pk, 0.05, 0.1, 4.0).cpu().numpy()
    uv0 = engine.fuse_uv2pt(p4, tab, dev(s["depths"]), 0.05, 0.1, 4.0).cpu().numpy()
    assert np.array_equal(uv, uv0)


def test_many_views_flush_paths_on_packed_frames(engine, scenes):
    """900 frames looking at the same points: byte counters flush mid-sweep (several flushes per warp), cells pass 255."""
    s = small_scene(scenes, orc, npoints=3000, nframes=3, width=96, height=72, seed=2)
    reps = 300
    wxyz, t = np.tile(s["wxyz"], (reps, 1)), np.tile(s["t"], (reps, 1))
    depths, masks = np.tile(s["depths"], (reps, 1, 1)), np.tile(s["masks"], (reps, 1, 1))
    ov1 = orc.fuse_project_vote(s["points"], s["K"], s["W"], s["H"], s["wxyz"], s["t"], s["depths"], s["masks"], 134, 0, 0.05, 0.1, 4.0, 4.0)
    ov = ov1 * reps
    assert ov.max() > 255
    tab = engine.FrameTable(s["K"], s["W"], s["H"], wxyz, t, 4.0)
    p4 = engine.pack_points(s["points"])
    for fmt in (3, 0):
        frames = engine.pack_frames(dev(depths), dev(masks), 3) if fmt == 3 else dev(depths)
        m = None if fmt == 3 else dev(masks)
        votes, labels = engine.fuse_project_vote_resolve(p4, tab, frames, m, 134, 133, 0.05, 0.1, 4.0, 0.5, None)
        assert np.array_equal(votes.cpu().numpy(), ov)
        assert np.array_equal(labels.cpu().numpy(), orc.segment(ov, 133, 0.5, None))
        _, l2 = engine.fuse_project_vote_resolve(p4, tab, frames, m, 134, 133, 0.05, 0.1, 4.0, 0.5, None, want_votes=False)   # uint16 histogram build
        assert np.array_equal(l2.cpu().numpy(), orc.segment(ov, 133, 0.5, None))


# ---------------------------------------------------------------------------------------------------------------------
# geometry kernels
# ---------------------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("case", ["cloud", "lattice"])
def test_radius_adjacency_matches_oracle_and_kdtree(engine, scenes, case):
    if case == "cloud":
        spec = scenes.scaled_spec("C1", npoints=20000, nframes=1, width=64, height=48, seed=4)
        p = scenes.make_cloud(spec).astype(np.float64)
        r = 0.1
    else:   # exact ties: lattice points at distance exactly r
        g = np.arange(12, dtype=np.float64) * 0.25
        p = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
        r = 0.5
    indptr, indices = engine.radius_adjacency(dev(p), r)
    oip, oix = orc.radius_adjacency(p, r)
    assert np.array_equal(indptr.cpu().numpy(), oip) and np.array_equal(indices.cpu().numpy(), oix)
    from sklearn.neighbors import KDTree
    rows = KDTree(p).query_radius(p[:500], r=r)                      # the reference's own call, fusion.py:374-375
    ip, ix = indptr.cpu().numpy(), indices.cpu().numpy()
    for i, row in enumerate(rows):
        assert np.array_equal(np.sort(row), ix[ip[i]:ip[i + 1]])


def test_split_into_instances_from_device_adjacency(engine, scenes):
    cv = importlib.import_module(PKG_NAME + ".Fusion3DSeg.segUtils.cv")
    spec = scenes.scaled_spec("C1", npoints=8000, nframes=1, width=64, height=48, seed=9)
    p = scenes.make_cloud(spec).astype(np.float64)
    rng = np.random.default_rng(0)
    classes = rng.integers(0, 5, len(p))
    classes[rng.random(len(p)) < 0.1] = 133
    oip, oix = orc.radius_adjacency(p, 0.12)
    want = orc.split_into_instances(classes, oip, oix, 133, None, 20)
    adj = engine.radius_adjacency(dev(p), 0.12)                       # stays on the device: no host CSR round trip
    got = cv.split_into_instances(classes, adj, 133, None, 20)
    assert len(got[0]) == want[0] and np.array_equal(got[1], want[1]) and got[2] == want[2] and np.array_equal(got[3], want[3])
    # one-directional adjacency lists (not the symmetric KDTree output): the reference BFS follows adj[point] as given
    keep = oix > np.repeat(np.arange(len(p)), np.diff(oip))           # only smaller -> larger entries
    ip2 = np.concatenate([[0], np.cumsum(np.add.reduceat(keep.astype(np.int64), oip[:-1]))])
    got2 = cv.split_into_instances(classes, (ip2, oix[keep]), 133, None, 20)
    assert np.array_equal(got2[1], want[1])


@pytest.mark.parametrize("model", ["pca", "aabb"])
def test_batched_obb_fit_matches_stated_box_model(engine, model):
    rng = np.random.default_rng(5)
    pts, ids = [], []
    for k in range(40):
        n = int(rng.integers(1, 400))
        yaw, pitch = rng.uniform(-1, 1, 2)
        R = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]]) @ \
            np.array([[1, 0, 0], [0, np.cos(pitch), -np.sin(pitch)], [0, np.sin(pitch), np.cos(pitch)]])
        pts.append(rng.uniform(-0.5, 0.5, (n, 3)) * np.array([2.0, 0.7, 0.2]) @ R.T + rng.uniform(-5, 5, 3))
        ids += [k] * n
    pts, ids = np.concatenate(pts), np.asarray(ids, dtype=np.int64)
    perm = rng.permutation(len(pts))
    pts, ids = pts[perm], ids[perm]
    want_ids = [3, 0, 17, 39, 8, 99]                                  # 99: no points
    boxes, counts = engine.obb_fit(dev(pts), dev(ids), want_ids, model)
    boxes, counts = boxes.cpu().numpy(), counts.cpu().numpy()
    for b, c, k in zip(boxes, counts, want_ids):
        sel = ids == k
        assert c == sel.sum()
        if c < 4:
            continue
        centre, R, extent = orc.fit_box(pts[sel], model)
        got = orc.box_corners(b[:3], b[3:12].reshape(3, 3), b[12:])
        ref = orc.box_corners(centre, R, extent)
        got, ref = got[np.lexsort(np.round(got, 6).T)], ref[np.lexsort(np.round(ref, 6).T)]
        np.testing.assert_allclose(got, ref, rtol=0, atol=1e-9)       # same box (axis signs may differ: compare corner sets)
        inside = engine.obb_contains(dev(pts), dev(b[None, :]))[0].cpu().numpy().astype(bool)
        assert (~inside[sel]).sum() <= 6                              # closed box of its own points; only the <= 6 face-defining points are ulp-fragile


def test_box_pairs_sweep_equals_brute_force_and_oracle(engine, scenes):
    lo, hi, group, _ = scenes.make_boxes(nboxes=6000, seed=3, extent=(20.0, 20.0, 4.0))
    lo[10], hi[10] = hi[10].copy(), lo[10].copy()                     # a degenerate (inverted) box
    lo[20:24, 0] = lo[20, 0]                                          # ties on lo.x
    sweep = engine.box_pairs_aabb(lo, hi, group).cpu().numpy().astype(np.int64)
    brute = engine.box_pairs_aabb(lo, hi, group, brute_force=True).cpu().numpy().astype(np.int64)
    a = np.unique(np.sort(sweep, axis=1), axis=0)
    b = np.unique(np.sort(brute, axis=1), axis=0)
    assert len(sweep) == len(a) and np.array_equal(a, b) and len(a) > 1000
    # reference predicate on every pair (merge_intersecting_bb.py:49-53)
    want = [(i, j) for i in range(300) for j in range(i + 1, 6000) if group[i] == group[j] and orc.aabb_overlap(lo[i], hi[i], lo[j], hi[j])]
    assert np.array_equal(a[a[:, 0] < 300], np.asarray(want, dtype=np.int64).reshape(-1, 2))


# ---------------------------------------------------------------------------------------------------------------------
# vote exchange: a dropped queue entry must raise
# ---------------------------------------------------------------------------------------------------------------------

def test_vote_exchange_overflow_is_caught(engine, scenes):
    import torch.distributed as dist
    parallel = importlib.import_module(PKG_NAME + ".parallel")
    if dist.is_initialized():
        pytest.skip("a process group already exists in this process")
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    try:
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    except Exception as ex:   # noqa: BLE001
        pytest.skip(f"no single-rank NCCL group: {ex}")
    try:
        s = small_scene(scenes, orc, npoints=20000, nframes=6, width=160, height=120, seed=7)
        ov = orc.fuse_project_vote(s["points"], s["K"], s["W"], s["H"], s["wxyz"], s["t"], s["depths"], s["masks"], 134, 0, 0.05, 0.1, 4.0, 4.0)
        tab = engine.FrameTable(s["K"], s["W"], s["H"], s["wxyz"], s["t"], 4.0)
        p4 = engine.pack_points(s["points"])
        pk = engine.pack_frames(dev(s["depths"]), dev(s["masks"]))

        def fuse(**xargs):
            engine.fuse_project_vote_exchange(p4, tab, pk, None, 134, radius=0.05, zmin=0.1, zmax=4.0, **xargs)
        try:
            good = parallel.VoteExchange(len(s["points"]), 134, torch.device("cuda", 0))
        except Exception as ex:   # noqa: BLE001
            pytest.skip(f"symmetric memory unavailable: {ex}")
        labels = good.run(fuse, 133, 0.5, None)
        assert np.array_equal(good.shard[:good.rows].cpu().numpy(), ov)
        assert np.array_equal(labels.cpu().numpy(), orc.segment(ov, 133, 0.5, None))
        # rows_per_block = 0 -> minimum record regions; sub_cap = 1 -> the first spilled / deferred entry overflows
        tiny = parallel.VoteExchange(len(s["points"]), 134, torch.device("cuda", 0), rows_per_block=0, sub_cap=1)
        tiny.sub_rows = 1
        with pytest.raises(RuntimeError, match="overflow"):
            tiny.run(fuse, 133, 0.5, None)
        tiny.run(fuse, 133, 0.5, None, check="deferred")              # deferred: the flag surfaces at finish()
        with pytest.raises(RuntimeError, match="overflow"):
            tiny.finish()
    finally:
        dist.destroy_process_group()


def test_spatquadranion_rotate_matches_oracle(engine):
    sq = importlib.import_module(PKG_NAME + ".RTAB_utils.spatQuad")
    rng = np.random.default_rng(1)
    p = rng.normal(size=(5000, 3)) * 3
    for q in ([0.9, 0.1, -0.3, 0.2], ["0.707107", "0", "0.707107", "0"], [2.0, 0.0, 0.0, 0.0]):
        Q = sq.SpatQuadranion(*q) if len(q) == 4 and isinstance(q[0], str) else sq.SpatQuadranion(q)
        qa = np.asarray([float(v) for v in q])
        assert np.array_equal(Q.rotate(p), orc.quat_rotate(qa, p))                   # un-normalised sandwich, bit for bit
        assert np.array_equal(Q.inverse.elements, orc.quat_inverse(qa))
        assert np.array_equal(Q.inverse.rotate(p), orc.quat_rotate(orc.quat_inverse(qa), p))


def test_unseen_supertiles_fast_path(engine, scenes):
    """Super-tiles no frame of the launch can see take the kernel's early exit: their vote rows must still be written
    (zeros over whatever the buffer held), labels = unclassified, and accumulation must leave them untouched."""
    s = small_scene(scenes, orc, npoints=9000, nframes=40, width=96, height=72, seed=13)     # > 32 frames: the first cull level runs
    far = np.random.default_rng(0).uniform(-1, 1, (3 * 4096, 3)).astype(np.float32) + np.float32(500.0)
    pts = np.concatenate([far[:4096], s["points"], far[4096:]])                              # unseen super-tiles before and after
    ov = orc.fuse_project_vote(pts, s["K"], s["W"], s["H"], s["wxyz"], s["t"], s["depths"], s["masks"], 134, 0, 0.05, 0.1, 4.0, 4.0)
    assert ov[:4096].sum() == 0 and ov[4096 + 9000:].sum() == 0 and ov.sum() > 1000
    tab = engine.FrameTable(s["K"], s["W"], s["H"], s["wxyz"], s["t"], 4.0)
    p4 = engine.pack_points(pts)
    pk = engine.pack_frames(dev(s["depths"]), dev(s["masks"]))
    for frames, m in ((pk, None), (dev(s["depths"]), dev(s["masks"]))):
        votes = torch.full((len(pts), 134), 77, dtype=torch.int32, device="cuda")
        labels = torch.full((len(pts),), -5, dtype=torch.int64, device="cuda")
        engine.fuse_project_vote_resolve(p4, tab, frames, m, 134, 133, 0.05, 0.1, 4.0, 0.5, None, votes=votes, labels=labels)
        assert np.array_equal(votes.cpu().numpy(), ov)
        assert np.array_equal(labels.cpu().numpy(), orc.segment(ov, 133, 0.5, None))
        engine.fuse_project_vote(p4, tab, frames, m, 134, 0.05, 0.1, 4.0, votes=votes, accumulate=True)
        assert np.array_equal(votes.cpu().numpy(), 2 * ov)
        v16 = torch.full((len(pts), 134), 9, dtype=torch.uint16, device="cuda")
        engine.fuse_project_vote(p4, tab, frames, m, 134, 0.05, 0.1, 4.0, votes=v16)
        assert np.array_equal(v16.cpu().numpy().astype(np.int64), ov)


def test_grouped_supertile_cull_keeps_every_contributing_frame(engine, scenes):
    """Large clouds take the grouped first-level cull (one CTA per 8 super-tiles).  A frame missing from a list would
    lose its votes; launches of <= 32 frames skip the first level altogether, so the chunked sum is the reference."""
    npts, nframes, W, H = 2368 * 4096 + 1234, 288, 64, 48          # just past the switch-over, ragged last super-tile
    spec = scenes.scaled_spec("C1", npoints=npts, nframes=nframes, width=W, height=H, seed=21)
    K = scenes.scaled_intrinsics(W, H)
    wxyz, t = scenes.make_poses(spec)
    pts = scenes.make_cloud(spec)
    rng = np.random.default_rng(5)
    depths = np.repeat(np.repeat(rng.integers(800, 3500, (nframes, H // 8, W // 8)), 8, axis=1), 8, axis=2).astype(np.uint16)
    masks = scenes.block_masks((H, W), nframes, seed=3, block=8)
    tab = engine.FrameTable(K, W, H, wxyz, t, 4.0)
    p4 = engine.pack_points(pts)
    pk = engine.pack_frames(dev(depths), dev(masks))
    one = engine.fuse_project_vote(p4, tab, pk, None, 134, 0.05, 0.1, 4.0)
    acc = None
    for a in range(0, nframes, 32):
        b = min(a + 32, nframes)
        acc = engine.fuse_project_vote(p4, tab, pk.slice(a, b), None, 134, 0.05, 0.1, 4.0, votes=acc, accumulate=acc is not None,
                                       frame_begin=a, frame_end=b)
    assert int(one.sum()) > 100000
    assert torch.equal(one, acc)
    # and against the oracle on a strided sample of the points
    idx = np.arange(0, npts, 4099)
    ov = orc.fuse_project_vote(pts[idx], K, W, H, wxyz, t, depths, masks, 134, 0, 0.05, 0.1, 4.0, 4.0)
    assert np.array_equal(one[torch.as_tensor(idx, device="cuda")].cpu().numpy(), ov)
