"""Multi-GPU label fusion: one process per GPU, frames sharded across ranks, votes combined with a reduce-scatter
over the point axis, labels resolved on each rank's shard and all-gathered.

Votes are integer sums over frames (`segUtils/voting.py:89-98`), so any frame partition gives bit-identical
results.  The point axis is processed in chunks: chunk k is fused on the compute stream while the reduce-scatter
of chunk k-1 runs on a communication stream (NCCL over NVLink / NVSwitch).  `combine_votes` works on any
backend (gloo has no reduce_scatter: all_reduce + slice) so the host logic is testable on CPU.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def frame_shard(nframes: int, rank: int, world: int):
    """Contiguous frame range [a, b) of `rank` (sizes differ by at most one)."""
    base, rem = divmod(nframes, world)
    a = rank * base + min(rank, rem)
    return a, a + base + (1 if rank < rem else 0)


def point_shard(npoints: int, rank: int, world: int):
    """Rows of a (padded) chunk owned by `rank` after the reduce-scatter: equal shares of ceil(n/world)."""
    per = -(-npoints // world)
    return min(rank * per, npoints), min((rank + 1) * per, npoints), per


def combine_votes(partial: torch.Tensor, group=None) -> torch.Tensor:
    """partial [n, C1] int32 on every rank -> this rank's [per, C1] slice of the element-wise sum (rows beyond n
    are zero padding)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n, c1 = partial.shape
    per = -(-n // world)
    if per * world != n:
        pad = torch.zeros((per * world - n, c1), dtype=partial.dtype, device=partial.device)
        partial = torch.cat([partial, pad], dim=0)
    if dist.get_backend(group) == "nccl":
        out = torch.empty((per, c1), dtype=partial.dtype, device=partial.device)
        dist.reduce_scatter_tensor(out, partial.contiguous(), op=dist.ReduceOp.SUM, group=group)
        return out
    full = partial.clone()
    dist.all_reduce(full, op=dist.ReduceOp.SUM, group=group)
    return full[rank * per:(rank + 1) * per].contiguous()


def gather_labels(shard: torch.Tensor, n: int, group=None) -> torch.Tensor:
    """Per-rank label slices [per] -> full [n] on every rank."""
    world = dist.get_world_size(group)
    out = torch.empty(shard.numel() * world, dtype=shard.dtype, device=shard.device)
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(out, shard.contiguous(), group=group)
    else:
        parts = [torch.empty_like(shard) for _ in range(world)]
        dist.all_gather(parts, shard.contiguous(), group=group)
        out = torch.cat(parts)
    return out[:n]


def fuse_sharded(fuse_chunk, resolve, npoints: int, nchunks: int, device, group=None):
    """Chunked pipeline.  `fuse_chunk(a, b)` -> partial votes [b-a, C1] of this rank's frames for points [a, b)
    (enqueued on the current stream); `resolve(votes)` -> labels [rows].  Returns full labels [npoints] on every
    rank.  On CUDA the collective of chunk k-1 overlaps the fusion of chunk k."""
    use_cuda = torch.device(device).type == "cuda"
    nchunks = max(1, min(nchunks, npoints)) if npoints else 1
    bounds = [(i * npoints // nchunks, (i + 1) * npoints // nchunks) for i in range(nchunks)]
    labels = torch.empty(npoints, dtype=torch.int64, device=device)
    if use_cuda:
        compute = torch.cuda.current_stream()
        comm = torch.cuda.Stream(device=device)
    keep = []
    for (a, b) in bounds:
        if b == a:
            continue
        part = fuse_chunk(a, b)
        if use_cuda:
            ev = torch.cuda.Event()
            ev.record(compute)
            with torch.cuda.stream(comm):
                comm.wait_event(ev)
                part.record_stream(comm)
                mine = combine_votes(part, group)
                lab = resolve(mine)
                full = gather_labels(lab, b - a, group)
                labels[a:b].copy_(full)
                keep.append((part, mine, lab, full))
        else:
            mine = combine_votes(part, group)
            labels[a:b] = gather_labels(resolve(mine), b - a, group)
    if use_cuda:
        done = torch.cuda.Event()
        done.record(comm)
        compute.wait_event(done)
        for tensors in keep:
            for t in tensors:
                t.record_stream(compute)
    return labels
