"""The fused level-P path as one host object: fixed cloud + per-frame poses / depth / masks -> votes -> labels.

This is the composition SURVEY 8(c) defines from reference lines (cull `fusion.py:254-260`, project `fusion.py:266`,
single-pixel criterion `fusion.py:223-228`, vote `voting.py:98`, resolve `voting.py:106-137`) executed by
`f3d_fuse_project_vote` + `f3d_resolve_labels`; frames may arrive in chunks (streaming ingest) and may be
sharded across ranks (`parallel.py`).
"""
from __future__ import annotations

import numpy as np
import torch

from . import engine
from ._lib import require_cuda


class FusedLabeler:
    def __init__(self, points, K, width, height, wxyz, translations, point_range=(0.1, 4.0), radius=0.05,
                 nclasses=133, max_depth=None):
        """points [N,3] (float32 values are the kernel contract; float64 input is rounded and `points_rounded`
        is set), K [3,3] scaled intrinsics, wxyz [F,4] (w,x,y,z) un-normalised, translations [F,3]
        (`parse_rts`, fusion.py:67-77); point_range / radius as `process3DSeg` (process3D.py:16-17);
        max_depth defaults to point_range[1] (process3D.py:39)."""
        require_cuda()
        if not isinstance(points, torch.Tensor):
            self.points_rounded = not engine.points_are_float32(points)
        else:
            self.points_rounded = points.dtype != torch.float32
        self.points4 = engine.pack_points(points)
        self.N = int(self.points4.shape[0])
        self.zmin, self.zmax = float(point_range[0]), float(point_range[1])
        self.radius = float(radius)
        self.nclasses = int(nclasses)
        self.table = engine.FrameTable(K, width, height, wxyz, translations,
                                       self.zmax if max_depth is None else max_depth)
        self.votes = None
        self.stats = engine.new_stats()

    @property
    def nframes(self):
        return self.table.F

    def vote(self, depths, masks, frame_begin=0, frame_end=None, accumulate=None, audit=False):
        """Fuse frames [frame_begin, frame_end) (device or host [F',H,W] depth uint16 mm / float32 m and uint8 masks).
        First call overwrites the vote tensor, later calls accumulate (unless `accumulate` says otherwise)."""
        depths = engine.as_cuda(depths)
        masks = engine.as_cuda(masks, torch.uint8)
        acc = (self.votes is not None) if accumulate is None else accumulate
        self.votes = engine.fuse_project_vote(self.points4, self.table, depths, masks, self.nclasses + 1, self.radius,
                                              self.zmin, self.zmax, votes=self.votes, accumulate=acc, stats=self.stats,
                                              audit=audit, frame_begin=frame_begin, frame_end=frame_end)
        return self.votes

    def segment(self, threshold=0.5, filter_classes=None, votes=None):
        """int64 [N] device labels (VotingSegmentation.segment semantics)."""
        v = self.votes if votes is None else votes
        return engine.resolve_labels(v, self.nclasses, threshold, filter_classes)

    def uv2pt(self, depths, frame_begin=0, frame_end=None):
        """The association in the reference's exchange format (fusion.py:253,297,322): int32 [F', H*W]."""
        return engine.fuse_uv2pt(self.points4, self.table, engine.as_cuda(depths), self.radius, self.zmin, self.zmax,
                                 stats=self.stats, frame_begin=frame_begin, frame_end=frame_end)

    def render_depth(self, border=0, frame_begin=0, frame_end=None):
        """Kernel (2): z-buffer splat of the cloud itself -> uint16 mm [F',H,W]."""
        return engine.zbuffer_splat(self.points4, self.table, border=border, frame_begin=frame_begin, frame_end=frame_end)

    def votes_numpy(self):
        """float64 [N, nclasses+1] -- the type `VotingSegmentation.votes` has in the reference (voting.py:34)."""
        return self.votes.to(torch.float64).cpu().numpy()

    def stats_dict(self):
        return engine.stats_dict(self.stats)


def fuse_labels(points, K, width, height, wxyz, translations, depths, masks, point_range=(0.1, 4.0), radius=0.05,
                nclasses=133, threshold=0.5, filter_classes=None, return_votes=True):
    """One-call public API used by bench.py's end-to-end measurement: host arrays in, host labels (and votes) out."""
    fl = FusedLabeler(points, K, width, height, wxyz, translations, point_range, radius, nclasses)
    fl.vote(depths, masks)
    labels = fl.segment(threshold, filter_classes).cpu().numpy()
    if return_votes:
        return fl.votes_numpy(), labels
    return labels


def _pin(a):
    """numpy / torch host array -> pinned torch tensor (no copy if already pinned)."""
    t = a if isinstance(a, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(a))
    return t if t.is_pinned() else t.pin_memory()


def vote_stream(fl: FusedLabeler, host_depths, host_masks, chunk_frames=64, frame_begin=0):
    """Streaming ingest: frames live in (ideally pinned) HOST memory and are copied to the device in chunks on a
    copy stream while the previous chunk is being fused on the compute stream (two staging buffers).  This is
    the end-to-end path bench.py times: host->device copies are inside the call."""
    F = int(host_depths.shape[0])
    dev = fl.points4.device
    compute = torch.cuda.current_stream()
    copy = torch.cuda.Stream(device=dev)
    hd = host_depths if isinstance(host_depths, torch.Tensor) else torch.as_tensor(host_depths)
    hm = host_masks if isinstance(host_masks, torch.Tensor) else torch.as_tensor(host_masks)
    nb = 2
    cf = max(1, min(chunk_frames, F))
    dbuf = [torch.empty((cf,) + tuple(hd.shape[1:]), dtype=hd.dtype, device=dev) for _ in range(nb)]
    mbuf = [torch.empty((cf,) + tuple(hm.shape[1:]), dtype=torch.uint8, device=dev) for _ in range(nb)]
    ready = [torch.cuda.Event() for _ in range(nb)]
    free = [torch.cuda.Event() for _ in range(nb)]
    for e in free:
        e.record(compute)
    k = 0
    for a in range(0, F, cf):
        b = min(a + cf, F)
        s = k % nb
        with torch.cuda.stream(copy):
            copy.wait_event(free[s])
            dbuf[s][: b - a].copy_(hd[a:b], non_blocking=True)
            mbuf[s][: b - a].copy_(hm[a:b], non_blocking=True)
            ready[s].record(copy)
        compute.wait_event(ready[s])
        fl.vote(dbuf[s][: b - a], mbuf[s][: b - a], frame_begin=frame_begin + a, frame_end=frame_begin + b)
        free[s].record(compute)
        k += 1
    return fl.votes


def fuse_labels_from_host(points, K, width, height, wxyz, translations, host_depths, host_masks,
                          point_range=(0.1, 4.0), radius=0.05, nclasses=133, threshold=0.5, filter_classes=None,
                          chunk_frames=64, out=None):
    """Public end-to-end call: everything starts in host memory, labels (int64 [N]) come back to host memory.
    Votes stay on the device (fetch them with `FusedLabeler.votes_numpy()` when needed)."""
    fl = FusedLabeler(points, K, width, height, wxyz, translations, point_range, radius, nclasses)
    vote_stream(fl, host_depths, host_masks, chunk_frames)
    labels = fl.segment(threshold, filter_classes)
    if out is None:
        out = torch.empty(labels.shape, dtype=labels.dtype, pin_memory=True)
    out.copy_(labels, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return out.numpy(), fl
