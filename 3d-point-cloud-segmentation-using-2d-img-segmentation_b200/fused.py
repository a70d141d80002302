"""The fused level-P path as one host object: fixed cloud + per-frame poses / depth / masks -> votes -> labels.

This is the composition SURVEY 8(c) defines from reference lines (cull `fusion.py:254-260`, project `fusion.py:266`,
single-pixel criterion `fusion.py:223-228`, vote `voting.py:98`, resolve `voting.py:106-137`) executed by
`f3d_fuse_project_vote` + `f3d_resolve_labels`; frames may arrive in chunks (streaming ingest) and may be
sharded across ranks (`parallel.py`).

Device layout of the frames: the reference's data contract is the FILES (16-bit depth PNGs `RTAB_utils/ios_rtab.py:97-113`,
uint8 mask PNGs `segUtils/voting.py:66`).  On the device the ingest path packs every pixel into one uint32 texel
(depth mm | class << 16, 16x16-pixel tiles: `engine.PackedFrames`, kernel `f3d_pack_frames` with the mask's nearest
resize `voting.py:93` folded in), because the fused sweep's frame gathers are bound by the NUMBER of 32-byte sectors they
touch, not by bytes: one sector per point-view instead of a depth and a mask sector.
"""
from __future__ import annotations

import numpy as np
import torch

from . import engine
from ._lib import FRAMES_U32_T16, require_cuda


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
        self.labels = None
        self.stats = engine.new_stats()
        self.frames = None            # PackedFrames of the whole scan once `ingest` / `pack` has run
        self._stage = None            # device staging buffers of the streaming ingest
        self._host_labels = None

    @property
    def nframes(self):
        return self.table.F

    @property
    def H(self):
        return self.table.H

    @property
    def W(self):
        return self.table.W

    # ---- frames ---------------------------------------------------------------------------------------------------
    def pack(self, depths, masks, frame_begin=0, fmt=FRAMES_U32_T16):
        """Pack device (or host) uint16-mm depth + uint8 masks (any mask resolution) of frames [frame_begin, ...) into
        the labeler's resident packed frame stack; returns the PackedFrames view of those frames."""
        depths = engine.as_cuda(depths)
        masks = engine.as_cuda(masks, torch.uint8)
        if self.frames is None or self.frames.fmt != fmt:
            self.frames = engine.PackedFrames.empty(self.nframes, self.H, self.W, fmt, self.points4.device)
        engine.pack_frames(depths, masks, out=self.frames, frame_begin=frame_begin)
        return self.frames.slice(frame_begin, frame_begin + int(depths.shape[0]))

    def vote(self, depths, masks=None, frame_begin=0, frame_end=None, accumulate=None, audit=False, pack=None):
        """Fuse frames [frame_begin, frame_end): `depths` = PackedFrames, or device / host [F',H,W] depth (uint16 mm /
        float32 m) with uint8 `masks`.  uint16 depth is packed first (`pack` defaults to True for it; masks of another
        resolution are nearest-resized like `voting.py:93`).  First call overwrites the vote tensor, later calls
        accumulate (unless `accumulate` says otherwise)."""
        if not isinstance(depths, engine.PackedFrames):
            depths = engine.as_cuda(depths)
            masks = engine.as_cuda(masks, torch.uint8)
            if pack is None:
                pack = depths.dtype == torch.uint16
            if pack:
                depths, masks = engine.pack_frames(depths, masks), None
            elif tuple(masks.shape[1:]) != (self.H, self.W):
                masks = engine.resize_nearest(masks, self.H, self.W)
        acc = (self.votes is not None) if accumulate is None else accumulate
        self.votes = engine.fuse_project_vote(self.points4, self.table, depths, masks, self.nclasses + 1, self.radius,
                                              self.zmin, self.zmax, votes=self.votes, accumulate=acc, stats=self.stats,
                                              audit=audit, frame_begin=frame_begin, frame_end=frame_end)
        return self.votes

    def segment(self, threshold=0.5, filter_classes=None, votes=None):
        """int64 [N] device labels (VotingSegmentation.segment semantics)."""
        v = self.votes if votes is None else votes
        return engine.resolve_labels(v, self.nclasses, threshold, filter_classes)

    def label(self, frames=None, threshold=0.5, filter_classes=None, want_votes=True, timer=None):
        """One launch over every frame with the label resolve fused into the kernel's epilogue (the vote tensor is never
        re-read).  `frames`: PackedFrames of all frames (default: the resident stack).  Returns device labels int64 [N];
        `self.votes` holds the votes when `want_votes`."""
        frames = self.frames if frames is None else frames
        if frames is None:
            raise ValueError("no frames: call pack() / ingest() first or pass PackedFrames")
        if self.labels is None:
            self.labels = torch.empty(self.N, dtype=torch.int64, device=self.points4.device)
        votes, labels = engine.fuse_project_vote_resolve(
            self.points4, self.table, frames, None, self.nclasses + 1, self.nclasses, self.radius, self.zmin, self.zmax,
            threshold, filter_classes, votes=self.votes if want_votes else None, want_votes=want_votes, labels=self.labels,
            stats=self.stats, timer=timer)
        if want_votes:
            self.votes = votes
        return labels

    def uv2pt(self, depths, frame_begin=0, frame_end=None):
        """The association in the reference's exchange format (fusion.py:253,297,322): int32 [F', H*W]."""
        if not isinstance(depths, engine.PackedFrames):
            depths = engine.as_cuda(depths)
        return engine.fuse_uv2pt(self.points4, self.table, depths, self.radius, self.zmin, self.zmax,
                                 stats=self.stats, frame_begin=frame_begin, frame_end=frame_end)

    def render_depth(self, border=0, frame_begin=0, frame_end=None):
        """Kernel (2): z-buffer splat of the cloud itself -> uint16 mm [F',H,W]."""
        return engine.zbuffer_splat(self.points4, self.table, border=border, frame_begin=frame_begin, frame_end=frame_end)

    def votes_numpy(self):
        """float64 [N, nclasses+1] -- the type `VotingSegmentation.votes` has in the reference (voting.py:34)."""
        return self.votes.to(torch.float64).cpu().numpy()

    def stats_dict(self):
        return engine.stats_dict(self.stats)

    # ---- streaming ingest (host frames -> resident packed stack) -------------------------------------------------------
    def ingest(self, host_depths, host_masks, chunk_frames=64, frame_begin=0):
        """Frames live in (ideally pinned) HOST memory: uint16 depth [F',H,W] and uint8 masks [F',mh,mw].  They are copied
        in chunks on a copy stream into two staging buffers and packed (+ mask resize) on the compute stream into the
        resident packed stack while the next chunk is in flight.  All buffers are allocated once per labeler."""
        F = int(host_depths.shape[0])
        dev = self.points4.device
        hd = host_depths if isinstance(host_depths, torch.Tensor) else torch.as_tensor(host_depths)
        hm = host_masks if isinstance(host_masks, torch.Tensor) else torch.as_tensor(host_masks)
        if hd.dtype != torch.uint16 or hm.dtype != torch.uint8:
            raise TypeError("ingest needs uint16 depth (mm) and uint8 masks")
        cf = max(1, min(chunk_frames, F))
        key = (cf, tuple(hd.shape[1:]), tuple(hm.shape[1:]))
        if self._stage is None or self._stage["key"] != key:
            self._stage = {
                "key": key, "copy": torch.cuda.Stream(device=dev),
                "d": [torch.empty((cf,) + tuple(hd.shape[1:]), dtype=torch.uint16, device=dev) for _ in range(2)],
                "m": [torch.empty((cf,) + tuple(hm.shape[1:]), dtype=torch.uint8, device=dev) for _ in range(2)],
                "ready": [torch.cuda.Event() for _ in range(2)], "free": [torch.cuda.Event() for _ in range(2)]}
        st = self._stage
        if self.frames is None:
            self.frames = engine.PackedFrames.empty(self.nframes, self.H, self.W, FRAMES_U32_T16, dev)
        compute = torch.cuda.current_stream()
        copy = st["copy"]
        for e in st["free"]:
            e.record(compute)
        k = 0
        for a in range(0, F, cf):
            b = min(a + cf, F)
            s = k & 1
            with torch.cuda.stream(copy):
                copy.wait_event(st["free"][s])
                st["d"][s][: b - a].copy_(hd[a:b], non_blocking=True)
                st["m"][s][: b - a].copy_(hm[a:b], non_blocking=True)
                st["ready"][s].record(copy)
            compute.wait_event(st["ready"][s])
            engine.pack_frames(st["d"][s][: b - a], st["m"][s][: b - a], out=self.frames, frame_begin=frame_begin + a)
            st["free"][s].record(compute)
            k += 1
        return self.frames

    def label_from_host(self, host_depths, host_masks, threshold=0.5, filter_classes=None, chunk_frames=64, host_points=None,
                        want_votes=True):
        """End-to-end step on a persistent labeler: (optionally the cloud,) depth and masks come from host memory, are
        packed into the resident stack, ONE fused launch produces votes + labels, the labels land in pinned host memory.
        Returns the host label array (int64 [N], valid on return)."""
        if host_points is not None:
            hp = host_points if isinstance(host_points, torch.Tensor) else torch.as_tensor(host_points)
            if hp.shape[1] == 4:
                self.points4.copy_(hp, non_blocking=True)
            else:
                self.points4[:, :3].copy_(hp, non_blocking=True)
        self.ingest(host_depths, host_masks, chunk_frames)
        labels = self.label(None, threshold, filter_classes, want_votes=want_votes)
        if self._host_labels is None:
            self._host_labels = torch.empty(self.N, dtype=torch.int64, pin_memory=True)
        self._host_labels.copy_(labels, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self._host_labels.numpy()


def fuse_labels(points, K, width, height, wxyz, translations, depths, masks, point_range=(0.1, 4.0), radius=0.05,
                nclasses=133, threshold=0.5, filter_classes=None, return_votes=True):
    """One-call API: host (or device) arrays in, host labels (and votes) out."""
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
    """Streaming ingest + vote of a frame range: host frames -> packed stack (copy stream overlapped with the packing) ->
    one accumulate launch over the range."""
    F = int(host_depths.shape[0])
    fl.ingest(host_depths, host_masks, chunk_frames, frame_begin)
    fl.vote(fl.frames.slice(frame_begin, frame_begin + F), None, frame_begin=frame_begin, frame_end=frame_begin + F)
    return fl.votes


def fuse_labels_from_host(points, K, width, height, wxyz, translations, host_depths, host_masks,
                          point_range=(0.1, 4.0), radius=0.05, nclasses=133, threshold=0.5, filter_classes=None,
                          chunk_frames=64, labeler=None):
    """Public end-to-end call: everything starts in host memory, labels (int64 [N]) come back to host memory.
    Pass the returned labeler back in (`labeler=`) to reuse its device buffers (vote tensor, packed frame stack, staging)
    for the next scan of the same shape.  Votes stay on the device (`labeler.votes_numpy()`)."""
    fl = labeler
    if fl is None:
        fl = FusedLabeler(points, K, width, height, wxyz, translations, point_range, radius, nclasses)
        host_points = None
    else:
        host_points = points
    out = fl.label_from_host(host_depths, host_masks, threshold, filter_classes, chunk_frames, host_points=host_points)
    return out, fl
