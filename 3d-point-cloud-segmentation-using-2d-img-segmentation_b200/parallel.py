"""Multi-GPU label fusion: one process per GPU, frames sharded across ranks, votes combined with a reduce-scatter
over the point axis, labels resolved on each rank's shard and all-gathered.

Votes are integer sums over frames (`segUtils/voting.py:89-98`), so any frame partition gives bit-identical
results.  The point axis is processed in chunks: chunk k is fused on the compute stream while the reduce-scatter /
resolve / all-gather of chunk k-1 runs on a communication stream (NCCL over NVLink / NVSwitch).  All buffers are
allocated once (`ShardedPipeline`), so a step issues no allocation and no host synchronisation.  With
`packed=True` partial votes travel as uint16 counters viewed as int32 pairs: the int32 sum never carries between the
halves while a vote tensor sees fewer than 65 536 frames, so the exchange is exact at half the bytes.
The collectives fall back to all_reduce / all_gather lists on gloo (no reduce_scatter there) so the host logic is
testable on CPU.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def init_process_group(local_rank: int):
    """NCCL process group whose internal stream is high priority (see ShardedPipeline)."""
    opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), pg_options=opts)


def frame_shard(nframes: int, rank: int, world: int):
    """Contiguous frame range [a, b) of `rank` (sizes differ by at most one)."""
    base, rem = divmod(nframes, world)
    a = rank * base + min(rank, rem)
    return a, a + base + (1 if rank < rem else 0)


def frame_shard_ids(nframes: int, rank: int, world: int, mode: str = "interleaved"):
    """Global frame indices of `rank`.  "contiguous" = frame_shard; "interleaved" = rank, rank + world, ...: on a loop
    trajectory a contiguous shard looks at 1/world of the cloud, which concentrates each rank's votes (and histogram
    flushes) on few tiles and its records on few owners; interleaving gives every rank the whole scene at 1/world of
    the frame rate.  Votes commute over frames, so both give identical results."""
    if mode == "contiguous":
        a, b = frame_shard(nframes, rank, world)
        return list(range(a, b))
    if mode != "interleaved":
        raise ValueError(mode)
    return list(range(rank, nframes, world))


def _reduce_scatter(out: torch.Tensor, full: torch.Tensor, group=None):
    """full [per*world, k] (sum over ranks) -> out [per, k] = this rank's slice."""
    if dist.get_backend(group) == "nccl":
        dist.reduce_scatter_tensor(out, full, op=dist.ReduceOp.SUM, group=group)
    else:
        tmp = full.clone()
        dist.all_reduce(tmp, op=dist.ReduceOp.SUM, group=group)
        r, per = dist.get_rank(group), out.shape[0]
        out.copy_(tmp[r * per:(r + 1) * per])


def _all_gather(out: torch.Tensor, shard: torch.Tensor, group=None):
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(out, shard, group=group)
    else:
        parts = [torch.empty_like(shard) for _ in range(dist.get_world_size(group))]
        dist.all_gather(parts, shard, group=group)
        out.copy_(torch.cat(parts))


class ShardedPipeline:
    """Chunked fuse -> reduce-scatter -> resolve -> all-gather pipeline with persistent buffers.

    `fuse_into(a, b, out)` must enqueue on the current stream the partial votes of this rank's frames for points
    [a, b) into `out[:b-a]` (shape [rows, c1], dtype `vote_dtype`); `resolve(votes, out_labels)` must enqueue the label
    resolve of `votes` [rows, c1] into `out_labels` [rows] int64."""

    def __init__(self, npoints: int, c1: int, nchunks: int, device, group=None, packed=True, total_frames=None):
        """`total_frames`: frames of ALL ranks (and all accumulate calls) that can vote into one tensor.  The packed uint16
        exchange is exact only while no cell can reach 65 536, so it is used only when total_frames is given and < 65 536;
        otherwise votes travel as int32."""
        self.group, self.device = group, torch.device(device)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.npoints, self.c1 = npoints, c1
        nchunks = max(1, min(nchunks, max(npoints, 1)))
        self.bounds = [(i * npoints // nchunks, (i + 1) * npoints // nchunks) for i in range(nchunks)]
        self.bounds = [(a, b) for a, b in self.bounds if b > a]
        big = max([b - a for a, b in self.bounds], default=0)
        self.per = -(-big // self.world) if big else 0
        self.cuda = self.device.type == "cuda"
        self.packed = bool(packed and self.cuda and c1 % 2 == 0 and total_frames is not None and 0 <= int(total_frames) < 65536)
        self.vote_dtype = torch.uint16 if self.packed else torch.int32
        rows = self.per * self.world
        self.part = [torch.zeros((rows, c1), dtype=self.vote_dtype, device=self.device) for _ in range(2)]
        self.mine = [torch.zeros((self.per, c1), dtype=self.vote_dtype, device=self.device) for _ in range(2)]
        self.lab = [torch.zeros(self.per, dtype=torch.int64, device=self.device) for _ in range(2)]
        self.full = [torch.zeros(rows, dtype=torch.int64, device=self.device) for _ in range(2)]
        self.labels = torch.empty(npoints, dtype=torch.int64, device=self.device)
        if self.cuda:
            # high priority: the exchange kernels (NCCL, resolve) are tiny next to a fuse launch that fills every SM and
            # must be scheduled ahead of its remaining CTAs, otherwise the exchange only starts when the sweep drains
            self.comm = torch.cuda.Stream(device=self.device, priority=-1)
            self.ready = [torch.cuda.Event() for _ in range(2)]
            self.free = [torch.cuda.Event() for _ in range(2)]

    def _exchange(self, s: int, a: int, b: int, resolve):
        n = b - a
        part, mine = self.part[s], self.mine[s]
        if self.packed:   # uint16 pairs summed as int32: exact, half the bytes
            _reduce_scatter(mine.view(torch.int32), part.view(torch.int32), self.group)
        else:
            _reduce_scatter(mine, part, self.group)
        resolve(mine, self.lab[s])
        _all_gather(self.full[s], self.lab[s], self.group)
        self.labels[a:b].copy_(self.full[s][:n])

    def run(self, fuse_into, resolve) -> torch.Tensor:
        """One pass over every chunk.  Returns the full label vector [npoints] (valid on the current stream)."""
        if not self.cuda:
            for (a, b) in self.bounds:
                self.part[0][b - a:].zero_()
                fuse_into(a, b, self.part[0])
                self._exchange(0, a, b, resolve)
            return self.labels
        compute = torch.cuda.current_stream(self.device)
        for e in self.free:
            e.record(compute)
        for k, (a, b) in enumerate(self.bounds):
            s = k & 1
            compute.wait_event(self.free[s])            # chunk k-2's exchange no longer reads part[s]
            if b - a < self.part[s].shape[0]:
                self.part[s][b - a:].zero_()            # padding rows (chunk not divisible by the world size)
            fuse_into(a, b, self.part[s])
            self.ready[s].record(compute)
            with torch.cuda.stream(self.comm):
                self.comm.wait_event(self.ready[s])
                self._exchange(s, a, b, resolve)
                self.free[s].record(self.comm)
        for e in self.free:
            compute.wait_event(e)
        return self.labels


def shard_points(npoints: int, world: int, tile: int = 256) -> int:
    """Points per owner rank of the record exchange: ceil(npoints / world) rounded up to the kernel's point tile."""
    per = -(-npoints // world)
    return max(tile, -(-per // tile) * tile)


class VoteExchange:
    """Frame-sharded fusion with the vote exchange fused into the compute kernel (CUDA only).

    Rank d owns the points [d*per, (d+1)*per), per a multiple of the 256-point tile.  Its receive buffer lives in
    symmetric memory (`torch.distributed._symmetric_memory`); per source rank it holds a record region (NREG
    sub-regions of variable-length records: per 32-point block L rows of 64 B, row j = the j-th class | count << 8 of
    each point), a directory {row offset, L} per block, NSUB (cell, count) sub-queues and their count table.  A step is
        barrier (peers are done reading the previous step) -> fused kernel: every warp writes its block's record and
        directory entry straight into the owner's memory over NVLink; deferred fp64 votes and spills go to the owner's
        sub-queues -> publish the queue counts -> barrier -> merge the G records of every owned point into the dense
        int32 shard row and the label -> apply the queue entries -> all-gather the labels.
    Nothing dense crosses the fabric and the sweep never writes a vote tensor; the owner writes its shard once."""

    def __init__(self, npoints: int, c1: int, device, group=None, rows_per_block=64, sub_cap=None, compact=None):
        import numpy as np
        import torch.distributed._symmetric_memory as symm
        from . import engine
        self.engine, self.np = engine, np
        self.group = dist.group.WORLD if group is None else group
        self.device = torch.device(device)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.npoints, self.c1 = npoints, c1
        G = self.world
        # compacted launches pay one stream synchronisation per step to skip the super-tiles this rank's frames cannot see:
        # worth it once the cloud is large (measured at 2 ranks, 20 M points, 100 frames each: sweep 0.54 -> 0.44 ms)
        self.compact = (G >= 2 and npoints >= (1 << 22)) if compact is None else bool(compact)
        self.nreg, self.nsub, self.nsub_fix, self.nlevel = engine.exchange_constants()
        self.per = shard_points(npoints, G)
        self.rows = max(0, min(self.per, npoints - self.rank * self.per))
        self.blocks = self.per // 32
        self.sub_rows = max(256, -(-self.blocks * int(rows_per_block) // self.nreg))     # 64-byte rows per record sub-region
        self.sub_cap = int(sub_cap) if sub_cap else max(512, -(-self.per // self.nsub))    # entries per sub-queue
        # int64 words: [queue: G x NSUB x sub_cap][counts: G x NSUB uint32][directory: G x blocks x NLEVEL][records: G x NREG x sub_rows x 8]
        self.q_words = G * self.nsub * self.sub_cap
        self.c_words = G * self.nsub // 2
        self.d_words = G * self.blocks * self.nlevel
        self.s_words = G * self.nreg * self.sub_rows * 8
        total = self.q_words + self.c_words + self.d_words + self.s_words
        self.rx = symm.empty(total, dtype=torch.int64, device=self.device)
        self.rx.zero_()
        self.hdl = symm.rendezvous(self.rx, self.group)
        base = [self.hdl.get_buffer(d, (total,), torch.int64).data_ptr() for d in range(G)]
        o_c, o_d, o_s = self.q_words, self.q_words + self.c_words, self.q_words + self.c_words + self.d_words
        r = self.rank
        self.peer_queue_ptrs = np.array([b + r * self.nsub * self.sub_cap * 8 for b in base], dtype=np.uint64)
        self.peer_count_ptrs = np.array([b + o_c * 8 for b in base], dtype=np.uint64)
        self.peer_dir_ptrs = np.array([b + (o_d + r * self.blocks * self.nlevel) * 8 for b in base], dtype=np.uint64)
        self.peer_slot_ptrs = np.array([b + (o_s + r * self.nreg * self.sub_rows * 8) * 8 for b in base], dtype=np.uint64)
        self.rx_queue, self.rx_count = self.rx[:self.q_words], self.rx[o_c:o_d]
        self.rx_dir, self.rx_slots = self.rx[o_d:o_s], self.rx[o_s:]
        self.cursors = torch.zeros(G * (self.nreg + self.nsub), dtype=torch.int32, device=self.device)
        self.overflow = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.ovf_any = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.ovf_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.ovf_event = torch.cuda.Event()
        self._ovf_pending = False
        self.shard = torch.zeros((max(self.per, 1), c1), dtype=torch.int32, device=self.device)
        self.lab = torch.zeros(max(self.per, 1), dtype=torch.int64, device=self.device)
        self.full = torch.zeros(max(self.per, 1) * G, dtype=torch.int64, device=self.device)
        # every rank's int16 copy of ALL labels, in symmetric memory: the owners store into the G copies from inside the
        # merge / queue-apply kernels (the label all-gather as peer stores under an HBM-bound kernel)
        self.full16 = symm.empty(max(self.per, 1) * G, dtype=torch.int16, device=self.device)
        self.full16.zero_()
        self.hdl16 = symm.rendezvous(self.full16, self.group)
        self.peer_label_ptrs = np.array([self.hdl16.get_buffer(d, (max(self.per, 1) * G,), torch.int16).data_ptr() for d in range(G)],
                                        dtype=np.uint64)

    def fuse_args(self):
        """Keyword arguments of engine.fuse_project_vote_exchange that describe this exchange."""
        return dict(nranks=self.world, points_per_shard=self.per, peer_slot_ptrs=self.peer_slot_ptrs,
                    peer_dir_ptrs=self.peer_dir_ptrs, peer_queue_ptrs=self.peer_queue_ptrs, sub_rows=self.sub_rows,
                    sub_cap=self.sub_cap, cursors=self.cursors, overflow=self.overflow, compact=self.compact)

    def run(self, fuse, nclasses_id, threshold=0.5, filter_classes=None, check=True, gather=True) -> torch.Tensor:
        """`fuse(**self.fuse_args())` enqueues the exchange-mode fused kernel over this rank's frames.  Returns labels
        [npoints]; `self.shard[:self.rows]` holds this rank's reduced votes.  `gather=False` skips the all-gather and
        returns only the labels of the points this rank owns, [rank*per, rank*per + rows) -- for callers that deliver
        each shard themselves (e.g. every rank copies its slice into one shared host array).

        A sub-queue that fills up DROPS entries (`xg_append`): the sender's flag is max-reduced over the ranks on the
        device right after the exchange and copied to pinned host memory.  `check=True` (default) waits for it and raises
        -- wrong labels never leave this call silently.  `check="deferred"` keeps the step free of host synchronisation
        (benchmark loops): the flag of step k is examined at the start of step k+1 and by `finish()`."""
        eng = self.engine
        self._raise_if_overflowed(wait=False)          # deferred flag of the previous step
        self.hdl.barrier(channel=0)
        self.cursors.zero_()
        self.overflow.zero_()
        fuse(**self.fuse_args())
        eng.exchange_publish(self.cursors, self.peer_count_ptrs, self.rank, self.sub_cap)
        self.hdl.barrier(channel=1)
        # labels are class ids < 2^15 (C1 <= 256 columns, filter values are columns): they travel as int16, stored by the
        # owner straight into every rank's copy while it merges; wider ids fall back to an int64 all-gather
        grp = self.group if self.group is not dist.group.WORLD else None
        narrow = gather and 0 <= int(nclasses_id) < 32768 and all(0 <= int(c) < 32768 for c in (filter_classes or ()))
        bc = dict(peer_labels16=self.peer_label_ptrs, first_point=self.rank * self.per) if narrow else {}
        if self.rows > 0:
            eng.exchange_merge(self.rx_slots, self.rx_dir, self.world, self.sub_rows, self.per, self.rows, self.c1, nclasses_id,
                               threshold, filter_classes, votes=self.shard, labels=self.lab, **bc)
            eng.exchange_queue_apply(self.rx_queue, self.rx_count, self.world, self.sub_cap, self.shard, self.rows, nclasses_id,
                                     self.lab, threshold, filter_classes, **bc)
        if narrow:
            self.hdl.barrier(channel=2)          # every owner's label stores have landed
            self.full.copy_(self.full16)
        elif gather:
            _all_gather(self.full, self.lab, grp)
        # overflow: any rank's dropped entry invalidates every rank's result
        self.ovf_any.copy_(self.overflow)
        if self.world > 1:
            dist.all_reduce(self.ovf_any, op=dist.ReduceOp.MAX, group=grp)
        self.ovf_host.copy_(self.ovf_any, non_blocking=True)
        self.ovf_event.record()
        self._ovf_pending = True
        if check is True:
            self._raise_if_overflowed(wait=True)
        if not gather:
            return self.lab[:self.rows]
        return self.full[:self.npoints]

    def _raise_if_overflowed(self, wait: bool):
        if not getattr(self, "_ovf_pending", False):
            return
        if wait:
            self.ovf_event.synchronize()
        elif not self.ovf_event.query():
            self.ovf_event.synchronize()   # the previous step has long finished when the next one starts
        self._ovf_pending = False
        if int(self.ovf_host[0]):
            raise RuntimeError("vote exchange: a (cell, count) sub-queue overflowed on some rank and votes were dropped; "
                               "the labels of this step are invalid -- enlarge sub_cap (VoteExchange(..., sub_cap=...)) or "
                               "use ShardedPipeline (dense reduce-scatter)")

    def finish(self):
        """Examine the overflow flag of the last `run(check="deferred")` step (synchronises)."""
        self._raise_if_overflowed(wait=True)

    def check_overflow(self):
        """Kept for callers of round 1: same as finish()."""
        self.finish()


def fuse_sharded(fuse_chunk, resolve, npoints: int, nchunks: int, device, group=None):
    """Convenience wrapper (allocates a pipeline per call): `fuse_chunk(a, b)` -> partial int32 votes [b-a, C1],
    `resolve(votes)` -> labels [rows]."""
    c1 = None

    def fuse_into(a, b, out):
        nonlocal c1
        v = fuse_chunk(a, b)
        out[:b - a].copy_(v)

    probe = fuse_chunk(0, min(1, npoints)) if npoints else None
    c1 = probe.shape[1] if probe is not None else 1
    pipe = ShardedPipeline(npoints, c1, nchunks, device, group, packed=False)
    return pipe.run(fuse_into, lambda v, out: out.copy_(resolve(v)))
