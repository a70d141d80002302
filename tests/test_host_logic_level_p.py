"""Level P drop-in (`FusedLabeler`, `Fusion.label_fixed_cloud` / `write_uv2pt_fixed_cloud` / `dump_data`, `process3DSeg`,
`fuse_labels`) WITHOUT a device.  The GPU operators these compose -- frame table, frame packing, fused project+vote(+resolve),
uv2pt, label resolve, radius adjacency -- are replaced by numpy stand-ins of the same call contracts built on the oracle, so
what runs here is the HOST logic: float32 rounding of the cloud, frame chunking and accumulation, the resident packed-frame
stack, file names and formats, returned dtypes.  Expected values are the vectors the unmodified reference produced (g2) or the
oracle.  The operators themselves are checked on the GPU (tests/test_gpu_*.py)."""
import importlib
import pickle
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from conftest import PKG_NAME, load_golden, small_scene
from oracle import f3d_oracle as orc

sys.path.insert(0, str(Path(__file__).parent / "golden"))
from make_golden_framedata import write_cache  # noqa: E402  (pure pickle writer)


class FakeTable:
    def __init__(self, K, width, height, wxyz, translations, max_depth):
        self.K = np.ascontiguousarray(np.asarray(K, dtype=np.float64).reshape(9))
        self.W, self.H, self.max_depth = int(width), int(height), float(max_depth)
        self.q = np.asarray(wxyz, dtype=np.float64).reshape(-1, 4)
        self.t = np.asarray(translations, dtype=np.float64).reshape(-1, 3)
        if len(self.q) != len(self.t):
            raise ValueError("wxyz and translations must have the same number of frames")
        self.F = len(self.t)
        self.table = torch.zeros(1, dtype=torch.uint8)

    def export(self):
        return tuple(torch.as_tensor(x) for x in orc.frustum_data(self.K.reshape(3, 3), self.W, self.H, self.q, self.t))


class FakePacked:
    """Stand-in of engine.PackedFrames: keeps the depth / class planes the uint32 texels would hold."""

    def __init__(self, depth, mask, fmt=3):
        self.depth, self.mask, self.fmt = depth, mask, int(fmt)
        self.H, self.W = depth.shape[1:]
        self.texels = depth

    @property
    def nframes(self):
        return len(self.depth)

    @classmethod
    def empty(cls, nframes, height, width, fmt=3, device=None):
        return cls(np.zeros((nframes, height, width), np.uint16), np.zeros((nframes, height, width), np.uint8), fmt)

    def slice(self, a, b):
        return FakePacked(self.depth[a:b], self.mask[a:b], self.fmt)


@pytest.fixture
def level_p(monkeypatch):
    eng = importlib.import_module(PKG_NAME + ".engine")
    fused = importlib.import_module(PKG_NAME + ".fused")
    cpu = torch.device("cpu")

    def planes(depth, mask):
        if isinstance(depth, FakePacked):
            return depth.depth, depth.mask, 0
        d = depth.numpy()
        return d, (None if mask is None else mask.numpy()), (0 if d.dtype == np.uint16 else 1)

    def pack_frames(depth, mask, fmt=3, out=None, frame_begin=0):
        if depth.dtype != torch.uint16 or mask.dtype != torch.uint8:
            raise TypeError("pack_frames needs uint16 depth (mm) and uint8 masks")
        F, H, W = depth.shape
        if out is None:
            out, frame_begin = FakePacked.empty(F, H, W, fmt), 0
        if (out.H, out.W) != (H, W) or frame_begin + F > out.nframes:
            raise ValueError("pack_frames: output stack does not match")
        out.depth[frame_begin:frame_begin + F] = depth.numpy()
        out.mask[frame_begin:frame_begin + F] = np.stack([orc.resize_nearest(m, W, H) for m in mask.numpy()]) if F else 0
        return out

    def oracle_votes(points4, table, depth, mask, c1, radius, zmin, zmax, fb, fe):
        d, m, fmt = planes(depth, mask)
        assert len(d) == fe - fb, "frames passed must cover [frame_begin, frame_end)"
        return orc.fuse_project_vote(points4[:, :3].double().numpy(), table.K.reshape(3, 3), table.W, table.H, table.q[fb:fe],
                                     table.t[fb:fe], d, m, c1, fmt, radius, zmin, zmax, table.max_depth)

    def fuse_project_vote(points4, table, depth, mask, nclasses1, radius=0.05, zmin=0.1, zmax=4.0, votes=None, accumulate=False,
                          stats=None, audit=False, frame_begin=0, frame_end=None, packed_u16=False, timer=None):
        fe = table.F if frame_end is None else frame_end
        ov = torch.as_tensor(oracle_votes(points4, table, depth, mask, nclasses1, radius, zmin, zmax, frame_begin, fe).astype(np.int32))
        if votes is None:
            return ov
        if accumulate:
            votes += ov
        else:
            votes.copy_(ov)
        return votes

    def fuse_project_vote_resolve(points4, table, depth, mask, nclasses1, nclasses_id, radius=0.05, zmin=0.1, zmax=4.0, threshold=0.5,
                                  filter_classes=None, votes=None, want_votes=True, labels=None, stats=None, audit=False,
                                  frame_begin=0, frame_end=None, timer=None):
        fe = table.F if frame_end is None else frame_end
        ov = oracle_votes(points4, table, depth, mask, nclasses1, radius, zmin, zmax, frame_begin, fe)
        lab = torch.as_tensor(orc.segment(ov, nclasses_id, threshold, filter_classes))
        if labels is not None:
            labels.copy_(lab)
            lab = labels
        if not want_votes:
            return None, lab
        ovt = torch.as_tensor(ov.astype(np.int32))
        if votes is not None:
            votes.copy_(ovt)
            ovt = votes
        return ovt, lab

    def fuse_uv2pt(points4, table, depth, radius=0.05, zmin=0.1, zmax=4.0, stats=None, audit=False, frame_begin=0, frame_end=None):
        fe = table.F if frame_end is None else frame_end
        d, _, fmt = planes(depth, None)
        assert len(d) == fe - frame_begin
        eyes, look, nrm = orc.frustum_data(table.K.reshape(3, 3), table.W, table.H, table.q, table.t)
        p = points4[:, :3].double().numpy()
        rows = [orc.frame_uv2pt(p, table.K.reshape(3, 3), table.W, table.H, table.q[f], table.t[f], eyes[f], look[f], nrm[f],
                                d[f - frame_begin], fmt, radius, zmin, zmax, table.max_depth) for f in range(frame_begin, fe)]
        return torch.as_tensor(np.stack(rows)) if rows else torch.zeros((0, table.H * table.W), dtype=torch.int32)

    for mod in (eng, fused):
        monkeypatch.setattr(mod, "require_cuda", lambda: cpu)
    monkeypatch.setattr(eng, "FrameTable", FakeTable)
    monkeypatch.setattr(eng, "PackedFrames", FakePacked)
    monkeypatch.setattr(eng, "pack_frames", pack_frames)
    monkeypatch.setattr(eng, "fuse_project_vote", fuse_project_vote)
    monkeypatch.setattr(eng, "fuse_project_vote_resolve", fuse_project_vote_resolve)
    monkeypatch.setattr(eng, "fuse_uv2pt", fuse_uv2pt)
    monkeypatch.setattr(eng, "resize_nearest", lambda masks, h, w: torch.as_tensor(np.stack([orc.resize_nearest(m, w, h) for m in masks.numpy()])))
    monkeypatch.setattr(eng, "resolve_labels", lambda votes, nid, thr=0.5, fc=None, out=None: torch.as_tensor(orc.segment(votes.numpy(), nid, thr, fc)))
    monkeypatch.setattr(eng, "radius_adjacency", lambda p, r: tuple(torch.as_tensor(np.asarray(x, dtype=np.int64)) for x in orc.radius_adjacency(p.numpy(), r)))
    return fused


def test_fusion_helpers_host_logic(level_p, tmp_path):
    fusion = importlib.import_module(PKG_NAME + ".Fusion3DSeg.fusion")
    g = load_golden("g2_levelp")
    W, H = int(g["W"]), int(g["H"])
    eyes, look, spokes, nrm = fusion.Fusion._get_frustum_data(g["K"], W, H, g["wxyz"], g["t"], frame_ids=[2, 0])
    oe, ol, on = orc.frustum_data(g["K"], W, H, g["wxyz"], g["t"])
    assert np.array_equal(eyes, oe[[2, 0]]) and np.array_equal(look, ol[[2, 0]]) and np.array_equal(nrm, on[[2, 0]])
    assert spokes.shape == (2, 4, 3) and np.array_equal(spokes[:, 3], oe[[2, 0]])
    votes, classes = fusion.Fusion.label_fixed_cloud(g["points"], g["K"], W, H, g["wxyz"], g["t"], g["depths"], g["masks"],
                                                     point_range=(0.1, 4), radius=0.05, filter_classes=[86, 114, 115])
    assert votes.dtype == np.float64 and np.array_equal(votes, g["votes"]) and np.array_equal(classes, g["seg_default"])
    names = [str(10 + i) for i in range(len(g["t"]))]
    out = fusion.Fusion.write_uv2pt_fixed_cloud(tmp_path, names, g["points"], g["K"], W, H, g["wxyz"], g["t"], g["depths"], chunk=3)
    assert sorted(p.name for p in out.iterdir()) == sorted(f"{n}.npy" for n in names)
    for i, n in enumerate(names):
        got = np.load(out / f"{n}.npy")
        assert got.dtype == np.int32 and np.array_equal(got, g["uv2pt"][i])          # reference-produced exchange files
    fusion.Fusion.dump_data(tmp_path, g["points"], nframes=len(names), depth_hw=(H, W), compute_adjacency=True, ds_radius=0.05)
    loaded = fusion.Fusion.load_data(tmp_path)
    assert len(loaded) == 8 and loaded[6] == (H, W) and np.array_equal(loaded[0], g["points"])
    oip, oix = orc.radius_adjacency(np.asarray(g["points"], dtype=np.float64), 0.1)
    assert len(loaded[7]) == len(g["points"]) and all(np.array_equal(loaded[7][i], oix[oip[i]:oip[i + 1]]) for i in range(0, len(oip) - 1, 53))


def test_fused_labeler_host_logic(level_p):
    g = load_golden("g2_levelp")
    W, H, F = int(g["W"]), int(g["H"]), len(g["t"])
    fl = level_p.FusedLabeler(g["points"], g["K"], W, H, g["wxyz"], g["t"], (0.1, 4), 0.05)
    assert fl.nframes == F and (fl.H, fl.W) == (H, W) and fl.points4.dtype == torch.float32 and tuple(fl.points4.shape) == (len(g["points"]), 4)
    # chunked accumulation: the first call overwrites, later calls add
    fl.vote(g["depths"][:2], g["masks"][:2], frame_begin=0, frame_end=2)
    fl.vote(g["depths"][2:], g["masks"][2:], frame_begin=2, frame_end=F)
    assert np.array_equal(fl.votes_numpy(), g["votes"])
    assert np.array_equal(fl.segment(0.5, [86, 114, 115]).numpy(), g["seg_default"])
    # resident packed stack filled in two pieces, then ONE fused call; persistent vote / label buffers are reused
    fl.pack(g["depths"][:3], g["masks"][:3], frame_begin=0)
    fl.pack(g["depths"][3:], g["masks"][3:], frame_begin=3)
    v_before, lab = fl.votes, fl.label(threshold=0.5, filter_classes=[86, 114, 115])
    assert fl.votes is v_before and np.array_equal(fl.votes_numpy(), g["votes"]) and np.array_equal(lab.numpy(), g["seg_default"])
    lab2 = fl.label(threshold=0.5, filter_classes=[86, 114, 115], want_votes=False)
    assert lab2 is lab and np.array_equal(lab2.numpy(), g["seg_default"])
    uv = fl.uv2pt(g["depths"][1:4], frame_begin=1, frame_end=4)
    assert np.array_equal(uv.numpy(), np.stack(g["uv2pt"][1:4]))
    with pytest.raises(ValueError):
        level_p.FusedLabeler(g["points"], g["K"], W, H, g["wxyz"], g["t"]).label()       # no frames yet
    votes, labels = level_p.fuse_labels(g["points"], g["K"], W, H, g["wxyz"], g["t"], g["depths"], g["masks"],
                                        filter_classes=[86, 114, 115])
    assert votes.dtype == np.float64 and np.array_equal(votes, g["votes"]) and np.array_equal(labels, g["seg_default"])
    # float64 coordinates that are not float32 values are rounded and flagged
    assert not fl.points_rounded
    assert level_p.FusedLabeler(np.asarray(g["points"], dtype=np.float64) + 1e-12, g["K"], W, H, g["wxyz"], g["t"]).points_rounded


def test_process3dseg_host_logic(level_p, scenes, tmp_path):
    p3d = importlib.import_module(PKG_NAME + ".Fusion3DSeg.process3D")
    s = small_scene(scenes, orc, npoints=4000, nframes=5, width=64, height=48, seed=21)
    F, h, w = s["depths"].shape
    frames = np.array([4, 7, 8, 15, 16])
    write_cache(tmp_path, s["depths"], frames)
    rts = {"intrinsic": s["K"], "intrinsicScaled": s["K"], "odo_wxyz": s["wxyz"][:, [1, 2, 3, 0]], "odo_xyz": s["t"], "RGB_res": (h, w, 3),
           "Depth_res": (h, w)}
    with open(tmp_path / "PointcloudMergeResults" / "rtscameradata_x.pkl", "wb") as fp:
        pickle.dump(rts, fp)
    out = tmp_path / "out"
    pts, norms, clrs, nmerges, occ, nframes, hw, adj = p3d.process3DSeg(str(tmp_path), str(out), radius=0.05, point_range=(0.1, 4),
                                                                        decimation=1, cloud=s["points"], chunk=2)
    assert nframes == F and tuple(hw) == (h, w) and np.array_equal(pts, s["points"].astype(np.float64))
    eyes, look, nrm = orc.frustum_data(s["K"], w, h, s["wxyz"], s["t"])
    p64 = s["points"].astype(np.float64)
    tot, seen = np.zeros(len(p64), np.int64), np.zeros(len(p64), np.int64)
    for f in range(F):
        want = orc.frame_uv2pt(p64, s["K"], w, h, s["wxyz"][f], s["t"][f], eyes[f], look[f], nrm[f], s["depths"][f], 0, 0.05, 0.1, 4.0, 4.0)
        got = np.load(out / "fusion" / "uv2pt" / f"{frames[f]}.npy")
        assert got.dtype == np.int32 and np.array_equal(got, want)
        hit = want[want >= 0]
        tot += np.bincount(hit, minlength=len(p64))
        seen[np.unique(hit)] += 1
    assert np.array_equal(nmerges, tot) and np.array_equal(occ, seen) and occ.dtype == np.uint32
    assert len(adj) == len(p64)
    # second run takes the cloud from fusion_data.pkl; decimation 2 leaves only the ::2 lattice matchable
    p3d.process3DSeg(str(tmp_path), str(out), radius=0.05, point_range=(0.1, 4), decimation=2)
    got = np.load(out / "fusion" / "uv2pt" / f"{frames[0]}.npy").reshape(h, w)
    lattice = np.zeros((h, w), bool)
    lattice[::2, ::2] = True
    assert (got[~lattice] == -1).all() and (got[lattice] >= 0).any()
    with pytest.raises(NotImplementedError):
        p3d.process3DSeg(str(tmp_path), str(tmp_path / "empty_out"))                     # no cloud given, none on disk
