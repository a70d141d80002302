"""PyTorch host plumbing over the libf3d C ABI: device buffers, streams and argument marshalling only -- all
arithmetic of the label-fusion path runs in the CUDA kernels of csrc/.  Every function takes / returns torch CUDA
tensors; the reference-shaped numpy API lives in the `Fusion3DSeg/` and `get3DSeg.py` mirrors next to this file.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import (DEPTH_F32_M, DEPTH_U16_MM, FRAMES_U32, FRAMES_U32_T16, NSTATS, STAT_NAMES, F3dError, check, host_f64, load, ptr,
                   require_cuda, stream_ptr)


def as_cuda(a, dtype=None) -> torch.Tensor:
    dev = require_cuda()
    if isinstance(a, torch.Tensor):
        t = a.to(device=dev, dtype=dtype) if (a.device != dev or (dtype is not None and a.dtype != dtype)) else a
    else:
        t = torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(dev)
    return t.contiguous()


def pack_points(points) -> torch.Tensor:
    """[N,3] (numpy / torch, any float dtype) -> float32 [N,4] device tensor (x, y, z, 0), the float4 stream the
    fused kernel reads.  The kernels' contract is float32 coordinates (exactly widened to fp64 in the exact path)."""
    p = as_cuda(points)
    if p.ndim != 2 or p.shape[1] not in (3, 4):
        raise ValueError("points must be [N,3]")
    if p.shape[1] == 4 and p.dtype == torch.float32:
        return p
    out = torch.zeros((p.shape[0], 4), dtype=torch.float32, device=p.device)
    out[:, :3] = p[:, :3].to(torch.float32)
    return out


def points_are_float32(points) -> bool:
    a = np.asarray(points)
    return a.dtype == np.float32 or bool(np.array_equal(a, a.astype(np.float32).astype(a.dtype)))


class FrameTable:
    """Device table of per-frame pose / frustum / projection records (f3d_frames_setup).
    Mirrors the state `Fusion.__init__` derives from `parse_rts` (Fusion3DSeg/fusion.py:94-103)."""

    def __init__(self, K, width, height, wxyz, translations, max_depth):
        load()
        dev = require_cuda()
        self.K = host_f64(K, 9)
        self.W, self.H = int(width), int(height)
        self.max_depth = float(max_depth)
        q = torch.as_tensor(np.ascontiguousarray(np.asarray(wxyz, dtype=np.float64).reshape(-1, 4))).to(dev)
        t = torch.as_tensor(np.ascontiguousarray(np.asarray(translations, dtype=np.float64).reshape(-1, 3))).to(dev)
        if len(q) != len(t):
            raise ValueError("wxyz and translations must have the same number of frames")
        self.F = int(len(t))
        nbytes = int(load().f3d_frame_table_bytes(max(self.F, 1)))
        self.table = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        if self.F:
            check(load().f3d_frames_setup(ptr(self.K), self.W, self.H, ptr(q), ptr(t), self.F, self.max_depth,
                                          ptr(self.table), stream_ptr()), "f3d_frames_setup")

    def export(self):
        dev = self.table.device
        eyes = torch.empty((self.F, 3), dtype=torch.float64, device=dev)
        look = torch.empty((self.F, 3), dtype=torch.float64, device=dev)
        nrm = torch.empty((self.F, 4, 3), dtype=torch.float64, device=dev)
        check(load().f3d_frames_export(ptr(self.table), self.F, ptr(eyes), ptr(look), ptr(nrm), stream_ptr()),
              "f3d_frames_export")
        return eyes, look, nrm


class PackedFrames:
    """Device-resident frame stack in the fused kernel's packed format: one uint32 texel per pixel = uint16 depth mm |
    class id << 16 (`f3d_pack_frames`), row-major or in 16x16-pixel tiles.  The on-disk contract (16-bit depth PNGs
    `RTAB_utils/ios_rtab.py:97-113`, uint8 mask PNGs `segUtils/voting.py:66`) is unchanged: this is the device layout
    the ingest path produces so that a point-view costs one 32-byte sector instead of two."""

    def __init__(self, texels: torch.Tensor, height: int, width: int, fmt: int = FRAMES_U32_T16):
        self.texels, self.H, self.W, self.fmt = texels, int(height), int(width), int(fmt)

    @property
    def nframes(self):
        return int(self.texels.shape[0])

    @staticmethod
    def texels_per_frame(height, width, fmt=FRAMES_U32_T16) -> int:
        return int(load().f3d_packed_frame_texels(int(height), int(width), int(fmt)))

    @classmethod
    def empty(cls, nframes, height, width, fmt=FRAMES_U32_T16, device=None):
        dev = require_cuda() if device is None else device
        t = torch.empty((int(nframes), cls.texels_per_frame(height, width, fmt)), dtype=torch.int32, device=dev)
        return cls(t, height, width, fmt)

    def slice(self, a, b):
        return PackedFrames(self.texels[a:b], self.H, self.W, self.fmt)


def pack_frames(depth, mask, fmt=FRAMES_U32_T16, out: PackedFrames | None = None, frame_begin=0) -> PackedFrames:
    """depth [F,H,W] uint16 mm + mask [F,mh,mw] uint8 (any resolution: nearest-resized with OpenCV's rule, voting.py:93)
    -> PackedFrames.  With `out`, frames are written at out[frame_begin : frame_begin + F] (streaming ingest)."""
    if depth.dtype != torch.uint16 or mask.dtype != torch.uint8:
        raise TypeError("pack_frames needs uint16 depth (mm) and uint8 masks")
    F, H, W = depth.shape
    if mask.shape[0] != F:
        raise ValueError("depth / mask frame counts differ")
    if out is None:
        out = PackedFrames.empty(F, H, W, fmt, depth.device)
        frame_begin = 0
    if (out.H, out.W) != (H, W) or frame_begin + F > out.nframes:
        raise ValueError("pack_frames: output stack does not match")
    if F:
        dst = out.texels[frame_begin:frame_begin + F]
        check(load().f3d_pack_frames(ptr(depth), ptr(mask), F, H, W, int(mask.shape[1]), int(mask.shape[2]), out.fmt, ptr(dst),
                                     stream_ptr()), "f3d_pack_frames")
    return out


def _frames_args(depth, mask, table, nf):
    """(depth pointer, format, mask pointer) of a frame stack given either as PackedFrames or as depth + mask tensors."""
    if isinstance(depth, PackedFrames):
        if nf > 0 and (depth.nframes != nf or (depth.H, depth.W) != (table.H, table.W)):
            raise ValueError("packed frames must be [frame_end-frame_begin] frames of the table's H x W")
        return (ptr(depth.texels) if nf else None), depth.fmt, None
    if nf > 0 and (depth.shape[0] != nf or tuple(depth.shape[1:]) != (table.H, table.W)
                   or (mask is not None and (mask.shape[0] != nf or tuple(mask.shape[1:]) != (table.H, table.W)))):
        raise ValueError("depth / mask must be [frame_end-frame_begin, H, W]")
    return (ptr(depth) if nf else None), (_depth_fmt(depth) if nf else 0), (ptr(mask) if (nf and mask is not None) else None)


def _depth_fmt(depth: torch.Tensor) -> int:
    if depth.dtype == torch.uint16:
        return DEPTH_U16_MM
    if depth.dtype == torch.float32:
        return DEPTH_F32_M
    raise TypeError("depth must be uint16 (millimetres) or float32 (metres)")


class KernelTimer:
    """CUDA events recorded by the library around the fused kernel alone (`f3d_fuse_time_next_call`): the events are
    the caller's, the library keeps nothing.  `arm()` before a fused call, `ms()` after synchronising."""

    def __init__(self):
        self.pairs = []

    def arm(self):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()   # torch creates the cudaEvent_t lazily: recording materialises the handle
        e1.record()
        check(load().f3d_fuse_time_next_call(e0.cuda_event, e1.cuda_event), "f3d_fuse_time_next_call")
        self.pairs.append((e0, e1))

    def ms(self):
        return np.array([a.elapsed_time(b) for a, b in self.pairs], dtype=np.float64)


_WORKSPACE = {}


def workspace(npoints: int, device) -> torch.Tensor:
    """Per-device scratch for the deferred fp64 fix-up queue (f3d_fuse_workspace_bytes); grown on demand, shared by
    calls on the same stream (calls on other streams should pass their own)."""
    need = int(load().f3d_fuse_workspace_bytes(int(npoints)))
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _WORKSPACE.get(key)
    if buf is None or buf.numel() < need:
        buf = torch.empty(need, dtype=torch.uint8, device=device)
        _WORKSPACE[key] = buf
    return buf


def new_stats() -> torch.Tensor:
    return torch.zeros(NSTATS, dtype=torch.int64, device=require_cuda())


def stats_dict(stats: torch.Tensor) -> dict:
    v = stats.cpu().tolist()
    return {n: int(v[i]) for i, n in enumerate(STAT_NAMES)}


def fuse_project_vote(points4, table: FrameTable, depth, mask, nclasses1, radius=0.05, zmin=0.1, zmax=4.0, votes=None,
                      accumulate=False, stats=None, audit=False, frame_begin=0, frame_end=None, packed_u16=False, timer=None):
    """Kernel (1).  depth / mask: [F', H, W] device tensors covering frames [frame_begin, frame_end), or depth =
    PackedFrames (mask ignored).  `votes` may be int32 (reference layout) or uint16 (packed exchange format, also
    selected by `packed_u16` when allocating)."""
    frame_end = table.F if frame_end is None else frame_end
    N = points4.shape[0]
    if votes is None:
        votes = torch.empty((N, nclasses1), dtype=torch.uint16 if packed_u16 else torch.int32, device=points4.device)
        accumulate = False
    nf = frame_end - frame_begin
    if votes.dtype == torch.uint16 and table.F >= 65536:
        raise ValueError("uint16 vote counters hold at most 65 535 frames per tensor: use int32 votes for this scan")
    dptr, fmt, mptr = _frames_args(depth, mask, table, nf)
    ws = workspace(N, points4.device)
    fn = load().f3d_fuse_project_vote_u16 if votes.dtype == torch.uint16 else load().f3d_fuse_project_vote
    if timer is not None:
        timer.arm()
    check(fn(
        ptr(points4), N, ptr(table.table), frame_begin, frame_end, dptr, fmt, mptr, table.H, table.W, ptr(table.K), float(radius),
        float(zmin), float(zmax), ptr(votes), int(nclasses1), int(bool(accumulate)), ptr(ws), ws.numel(), ptr(stats),
        int(bool(audit)), stream_ptr()), "f3d_fuse_project_vote")
    return votes


def fuse_project_vote_resolve(points4, table: FrameTable, depth, mask, nclasses1, nclasses_id, radius=0.05, zmin=0.1,
                              zmax=4.0, threshold=0.5, filter_classes=None, votes=None, want_votes=True, labels=None,
                              stats=None, audit=False, frame_begin=0, frame_end=None, timer=None):
    """Kernel (1) with the label resolve fused into its epilogue.  Returns (votes or None, labels int64 [N]).
    `timer`: a KernelTimer that gets the duration of the fused kernel alone."""
    frame_end = table.F if frame_end is None else frame_end
    N = points4.shape[0]
    if votes is None and want_votes:
        votes = torch.empty((N, nclasses1), dtype=torch.int32, device=points4.device)
    if labels is None:
        labels = torch.empty(N, dtype=torch.int64, device=points4.device)
    filt = None if filter_classes is None else np.ascontiguousarray(np.asarray(filter_classes, dtype=np.int32))
    nf = frame_end - frame_begin
    dptr, fmt, mptr = _frames_args(depth, mask, table, nf)
    ws = workspace(N, points4.device)
    if timer is not None:
        timer.arm()
    check(load().f3d_fuse_project_vote_resolve(
        ptr(points4), N, ptr(table.table), frame_begin, frame_end, dptr, fmt, mptr, table.H, table.W, ptr(table.K), float(radius),
        float(zmin), float(zmax), ptr(votes), int(nclasses1), float(threshold), ptr(filt),
        0 if filt is None else int(filt.size), int(nclasses_id), ptr(labels), ptr(ws), ws.numel(), ptr(stats),
        int(bool(audit)), stream_ptr()), "f3d_fuse_project_vote_resolve")
    return votes, labels


def _filter_arg(filter_classes):
    filt = None if filter_classes is None else np.ascontiguousarray(np.asarray(filter_classes, dtype=np.int32))
    if filt is not None and filt.size == 0:
        raise ValueError("filter_classes must not be empty")
    return filt, (0 if filt is None else int(filt.size))


def exchange_constants():
    """(NREG, NSUB, NSUB_FIX, NLEVEL): record sub-regions, sub-queues, fix-up-owned sub-queues per (source, owner) and
    directory levels per 32-point block."""
    out = np.zeros(4, dtype=np.int32)
    check(load().f3d_exchange_constants(ptr(out)), "f3d_exchange_constants")
    return tuple(int(v) for v in out)


def fuse_project_vote_exchange(points4, table: FrameTable, depth, mask, nclasses1, nranks, points_per_shard, peer_slot_ptrs,
                               peer_dir_ptrs, peer_queue_ptrs, sub_rows, sub_cap, cursors, overflow, radius=0.05, zmin=0.1,
                               zmax=4.0, stats=None, frame_begin=0, frame_end=None, timer=None, compact=False):
    """Kernel (1) with the multi-GPU exchange fused in: the votes of this rank's frames go straight into the owner
    ranks' memory through the peer pointers (numpy uint64 [G] each) as slot records + directory entries, and as
    (cell, count) queue entries for what does not go into a record.  `cursors` is uint32 [G * (NREG + NSUB)], zeroed
    by the caller.  No dense vote tensor is written.  `compact`: launch only over the super-tiles this rank's frames can
    see (F3D_FUSE_COMPACT: one stream synchronisation inside the call, same results)."""
    frame_end = table.F if frame_end is None else frame_end
    N = points4.shape[0]
    ws = workspace(N, points4.device)
    arrs = [np.ascontiguousarray(np.asarray(a, dtype=np.uint64)) for a in (peer_slot_ptrs, peer_dir_ptrs, peer_queue_ptrs)]
    if any(a.size != nranks for a in arrs):
        raise ValueError("peer pointer arrays must have one entry per rank")
    nreg, nsub = exchange_constants()[:2]
    if cursors.dtype != torch.int32 or cursors.numel() < nranks * (nreg + nsub):
        raise ValueError("cursors must be int32 [nranks * (NREG + NSUB)]")
    dptr, fmt, mptr = _frames_args(depth, mask, table, frame_end - frame_begin)
    if timer is not None:
        timer.arm()
    check(load().f3d_fuse_project_vote_exchange(
        ptr(points4), N, ptr(table.table), frame_begin, frame_end, dptr, fmt, mptr, table.H, table.W,
        ptr(table.K), float(radius), float(zmin), float(zmax), int(nclasses1), int(nranks), int(points_per_shard), ptr(arrs[0]),
        ptr(arrs[1]), ptr(arrs[2]), int(sub_rows), int(sub_cap), ptr(cursors), ptr(overflow), ptr(ws), ws.numel(), ptr(stats),
        2 if compact else 0, stream_ptr()), "f3d_fuse_project_vote_exchange")


def exchange_publish(cursors, peer_count_ptrs, rank, sub_cap):
    c = np.ascontiguousarray(np.asarray(peer_count_ptrs, dtype=np.uint64))
    check(load().f3d_exchange_publish(ptr(cursors), ptr(c), int(rank), int(c.size), int(sub_cap), stream_ptr()),
          "f3d_exchange_publish")


def exchange_merge(slots, dirs, nranks, sub_rows, points_per_shard, nrows, nclasses1, nclasses_id, threshold=0.5,
                   filter_classes=None, votes=None, labels=None, peer_labels16=None, first_point=0):
    """Owner side: merge the slot records of all source ranks into the dense int32 shard rows and the labels.
    `peer_labels16` (numpy uint64 [nranks] of device pointers to every rank's int16 label array of all points) makes the
    kernel store each label at `first_point + row` of all of them -- the label all-gather rides inside the merge."""
    filt, nf = _filter_arg(filter_classes)
    pl = None if peer_labels16 is None else peer_labels16.ctypes.data
    check(load().f3d_exchange_merge(ptr(slots), ptr(dirs), int(nranks), int(sub_rows), int(points_per_shard), int(nrows),
                                    int(nclasses1), float(threshold), ptr(filt), nf, int(nclasses_id), ptr(votes), ptr(labels),
                                    pl, int(first_point), stream_ptr()), "f3d_exchange_merge")


def exchange_queue_apply(queue, counts, nranks, sub_cap, votes, nrows, nclasses_id, labels=None, threshold=0.5,
                         filter_classes=None, peer_labels16=None, first_point=0):
    """Owner side: scatter-add the received (cell, count) entries into the shard and re-resolve the touched points
    (`peer_labels16` / `first_point` as in exchange_merge)."""
    filt, nf = _filter_arg(filter_classes)
    pl = None if peer_labels16 is None else peer_labels16.ctypes.data
    check(load().f3d_exchange_queue_apply(ptr(queue), ptr(counts), int(nranks), int(sub_cap), ptr(votes), int(nrows),
                                          votes.shape[1], float(threshold), ptr(filt), nf, int(nclasses_id), ptr(labels),
                                          pl, int(first_point), stream_ptr()), "f3d_exchange_queue_apply")


def fuse_uv2pt(points4, table: FrameTable, depth, radius=0.05, zmin=0.1, zmax=4.0, stats=None, audit=False,
               frame_begin=0, frame_end=None):
    frame_end = table.F if frame_end is None else frame_end
    nf = frame_end - frame_begin
    uv2pt = torch.full((nf, table.H * table.W), -1, dtype=torch.int32, device=points4.device)
    if nf:
        ws = workspace(points4.shape[0], points4.device)
        dptr, fmt, _ = _frames_args(depth, None, table, nf)
        check(load().f3d_fuse_uv2pt(ptr(points4), points4.shape[0], ptr(table.table), frame_begin, frame_end, dptr,
                                    fmt, table.H, table.W, ptr(table.K), float(radius), float(zmin),
                                    float(zmax), ptr(uv2pt), ptr(ws), ws.numel(), ptr(stats), int(bool(audit)),
                                    stream_ptr()), "f3d_fuse_uv2pt")
    return uv2pt


def zbuffer_splat(points4, table: FrameTable, border=0, stats=None, audit=False, frame_begin=0, frame_end=None,
                  zbuf=None, out=None):
    """Kernel (2): uint16-millimetre depth images [F', H, W] of the cloud seen from frames [frame_begin, frame_end)."""
    frame_end = table.F if frame_end is None else frame_end
    nf = frame_end - frame_begin
    dev = points4.device
    if zbuf is None:
        zbuf = torch.empty((nf, table.H * table.W), dtype=torch.int32, device=dev)
    if out is None:
        out = torch.empty((nf, table.H, table.W), dtype=torch.uint16, device=dev)
    if nf:
        ws = workspace(points4.shape[0], points4.device)
        check(load().f3d_zbuffer_splat(ptr(points4), points4.shape[0], ptr(table.table), frame_begin, frame_end, table.H,
                                       table.W, ptr(table.K), ptr(zbuf), ptr(out), int(border), ptr(ws), ws.numel(),
                                       ptr(stats), int(bool(audit)), stream_ptr()), "f3d_zbuffer_splat")
    return out


def vote_uv2pt(votes_packed, uv2pt, mask, first_tag):
    """Level V: uv2pt [F, HW] int32, mask [F, HW] uint8 (already at depth resolution), votes_packed [N, C1] int32."""
    F, npix = uv2pt.shape
    N, C1 = votes_packed.shape
    check(load().f3d_vote_uv2pt(ptr(uv2pt), ptr(mask), F, npix, int(first_tag), ptr(votes_packed), N, C1, stream_ptr()),
          "f3d_vote_uv2pt")
    return votes_packed


def vote_finalize(votes_packed):
    check(load().f3d_vote_finalize(ptr(votes_packed), votes_packed.numel(), stream_ptr()), "f3d_vote_finalize")
    return votes_packed


def resize_nearest(masks, height, width):
    """[n, sh, sw] uint8 -> [n, height, width] with OpenCV's INTER_NEAREST index rule."""
    n, sh, sw = masks.shape
    out = torch.empty((n, height, width), dtype=torch.uint8, device=masks.device)
    check(load().f3d_resize_nearest_u8(ptr(masks), n, sh, sw, ptr(out), height, width, stream_ptr()),
          "f3d_resize_nearest_u8")
    return out


def resolve_labels(votes, nclasses_id, threshold=0.5, filter_classes=None, out=None):
    """Kernel (3): votes [N, C1] int32 -> int64 [N]."""
    N, C1 = votes.shape
    if out is None:
        out = torch.empty(N, dtype=torch.int64, device=votes.device)
    filt = None if filter_classes is None else np.ascontiguousarray(np.asarray(filter_classes, dtype=np.int32))
    if filt is not None and filt.size == 0:
        raise ValueError("filter_classes must not be empty")
    fn = load().f3d_resolve_labels_u16 if votes.dtype == torch.uint16 else load().f3d_resolve_labels
    check(fn(ptr(votes), N, C1, float(threshold), ptr(filt), 0 if filt is None else int(filt.size),
                                    int(nclasses_id), ptr(out), stream_ptr()), "f3d_resolve_labels")
    return out


def project_pixels(points64, K, quat, translation):
    p = as_cuda(points64, torch.float64)
    N = p.shape[0]
    uv = torch.empty((2, N), dtype=torch.int32, device=p.device)
    check(load().f3d_project_pixels(ptr(p), N, ptr(host_f64(K, 9)), ptr(host_f64(quat, 4)), ptr(host_f64(translation, 3)),
                                    ptr(uv), stream_ptr()), "f3d_project_pixels")
    return uv


def quat_rotate(points64, wxyz):
    """`SpatQuadranion(wxyz).rotate(points)` (`RTAB_utils/spatQuad.py:6-28`) -> float64 [N,3] device tensor."""
    p = as_cuda(points64, torch.float64)
    out = torch.empty_like(p)
    check(load().f3d_quat_rotate(ptr(p), p.shape[0], ptr(host_f64(wxyz, 4)), ptr(out), stream_ptr()), "f3d_quat_rotate")
    return out


def frustum_mask(points64, plane_points, normals):
    p = as_cuda(points64, torch.float64)
    N = p.shape[0]
    pp, nn = host_f64(plane_points), host_f64(normals)
    if pp.size != nn.size or pp.size % 3:
        raise ValueError("plane_points / normals must both be [M,3]")
    out = torch.empty(N, dtype=torch.uint8, device=p.device)
    check(load().f3d_frustum_mask(ptr(p), N, ptr(pp), ptr(nn), pp.size // 3, ptr(out), stream_ptr()), "f3d_frustum_mask")
    return out


def box_pairs_aabb(lo, hi, group, cap=None, brute_force=False):
    """All i<j pairs of equal group whose closed AABBs overlap (`merge_intersecting_bb.py:49-53`).  -> int32 [E,2]
    (unordered).  Default: sort-and-sweep broad phase over the boxes ordered by (group, lo.x); `brute_force=True` runs the
    all-pairs tile kernel (kept as the cross-check)."""
    lo, hi = as_cuda(lo, torch.float64), as_cuda(hi, torch.float64)
    group = as_cuda(group, torch.int32)
    B = lo.shape[0]
    cap = max(1024, 8 * B) if cap is None else cap
    order = None
    if not brute_force and B:
        o1 = torch.sort(lo[:, 0].contiguous(), stable=True).indices            # by lo.x ...
        order = o1[torch.sort(group[o1], stable=True).indices].to(torch.int32).contiguous()   # ... then (stable) by group
    while True:
        edges = torch.empty((cap, 2), dtype=torch.int32, device=lo.device)
        count = torch.zeros(1, dtype=torch.int64, device=lo.device)
        if order is None:
            check(load().f3d_box_pairs_aabb(ptr(lo), ptr(hi), ptr(group), B, ptr(edges), cap, ptr(count), stream_ptr()),
                  "f3d_box_pairs_aabb")
        else:
            check(load().f3d_box_pairs_sweep(ptr(lo), ptr(hi), ptr(group), ptr(order), B, ptr(edges), cap, ptr(count), stream_ptr()),
                  "f3d_box_pairs_sweep")
        n = int(count.item())
        if n <= cap:
            return edges[:n]
        cap = n


OBB_MODELS = {"pca": 0, "aabb": 1}


def obb_fit(points64, ids, instance_ids, model="pca"):
    """Boxes of the listed instances in ONE pass over the cloud (`f3d_obb_fit`).  points64 [N,3] float64 device tensor,
    ids [N] int64 device tensor, instance_ids: sequence of id values.  -> (boxes [L,15] float64: centre, R row major,
    extent; counts [L] int64), device tensors."""
    inst = np.asarray(instance_ids, dtype=np.int64).reshape(-1)
    L = int(inst.size)
    dev = points64.device
    boxes = torch.zeros((L, 15), dtype=torch.float64, device=dev)
    counts = torch.zeros(L, dtype=torch.int64, device=dev)
    if L == 0:
        return boxes, counts
    if inst.min() < 0:
        raise ValueError("instance ids must be non-negative")
    nslot = int(inst.max()) + 1
    table = np.full(nslot, -1, dtype=np.int32)
    table[inst[::-1]] = np.arange(L - 1, -1, -1, dtype=np.int32)           # first occurrence wins
    slot = torch.as_tensor(table).to(dev)
    ws = torch.empty(int(load().f3d_obb_fit_workspace_bytes(L)), dtype=torch.uint8, device=dev)
    check(load().f3d_obb_fit(ptr(points64), ptr(ids), int(points64.shape[0]), ptr(slot), nslot, L, OBB_MODELS[model], ptr(boxes),
                             ptr(counts), ptr(ws), stream_ptr()), "f3d_obb_fit")
    return boxes, counts


def radius_adjacency(points64, r):
    """`KDTree(points).query_radius(points, r)` (`fusion.py:374-375`) as a CSR pair of device tensors
    (indptr int64 [N+1], indices int64), rows sorted ascending.  Uniform grid of cell size r; torch provides the key sort
    and the exclusive scan (plumbing), the kernels of csrc/geometry.cu the cell keys, the counting and the fill."""
    p = as_cuda(points64, torch.float64)
    N = int(p.shape[0])
    dev = p.device
    if N == 0:
        return torch.zeros(1, dtype=torch.int64, device=dev), torch.zeros(0, dtype=torch.int64, device=dev)
    r = float(r)
    mn = host_f64(p.min(0).values.cpu().numpy(), 3)
    mx = host_f64(p.max(0).values.cpu().numpy(), 3)
    keys = torch.empty(N, dtype=torch.int64, device=dev)
    check(load().f3d_radius_grid_keys(ptr(p), N, ptr(mn), ptr(mx), r, ptr(keys), stream_ptr()), "f3d_radius_grid_keys")
    skeys, order = torch.sort(keys, stable=True)
    counts = torch.empty(N, dtype=torch.int64, device=dev)
    check(load().f3d_radius_adjacency(ptr(p), N, ptr(mn), ptr(mx), r, ptr(skeys), ptr(order), ptr(counts), None, None, stream_ptr()),
          "f3d_radius_adjacency(count)")
    indptr = torch.zeros(N + 1, dtype=torch.int64, device=dev)
    torch.cumsum(counts, 0, out=indptr[1:])
    indices = torch.empty(int(indptr[-1].item()), dtype=torch.int64, device=dev)
    check(load().f3d_radius_adjacency(ptr(p), N, ptr(mn), ptr(mx), r, ptr(skeys), ptr(order), None, ptr(indptr), ptr(indices), stream_ptr()),
          "f3d_radius_adjacency(fill)")
    return indptr, indices


def union_find(nboxes, edges):
    labels = torch.empty(nboxes, dtype=torch.int32, device=require_cuda())
    E = 0 if edges is None else int(edges.shape[0])
    check(load().f3d_union_find(int(nboxes), ptr(edges) if E else None, E, ptr(labels), stream_ptr()), "f3d_union_find")
    return labels


def obb_contains(points64, boxes15):
    p = as_cuda(points64, torch.float64)
    b = as_cuda(boxes15, torch.float64).reshape(-1, 15)
    out = torch.empty((b.shape[0], p.shape[0]), dtype=torch.uint8, device=p.device)
    check(load().f3d_obb_contains(ptr(p), p.shape[0], ptr(b), b.shape[0], ptr(out), stream_ptr()), "f3d_obb_contains")
    return out
