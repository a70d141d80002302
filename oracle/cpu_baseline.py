"""CPU arm of the benchmark -- TEST / MEASUREMENT INFRASTRUCTURE, NOT PRODUCT CODE.

Times the numpy restatement of the reference's label-fusion path (`oracle/f3d_oracle.py`, kind = "port": the
reference itself is Python + third-party modules that are absent from this image and `/root/reference` does not
exist on the GPU box) on the host cores, the way the reference times itself (`time.perf_counter` around the stage,
`get3DSeg.py:75,85`).  Two figures:

  * as shipped -- ONE process, frames one after the other (`CpuFusion(workers=1)`): the reference has no parallelism;
  * frame-sharded over all host cores (`multiprocessing`, spawn): every worker holds the cloud AND the sample's
    depth / mask images resident (attached once from `multiprocessing.shared_memory`), so a pass ships only frame
    indices to the workers and sparse (point, class) hits back; the parent adds them with `votes[idx, cls] += 1`
    (`segUtils/voting.py:98`), and `segment` (`voting.py:106-137`) is row-sharded over the same workers on a vote
    tensor that also lives in shared memory.

Only `bench.py` (cpu_baseline leg and `--impl reference`) imports this module.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time
from multiprocessing import shared_memory

import numpy as np

from . import f3d_oracle as orc

_G = {}


def _attach(desc):
    name, shape, dtype = desc
    shm = shared_memory.SharedMemory(name=name)
    return shm, np.ndarray(shape, dtype=np.dtype(dtype), buffer=shm.buf)


def _init(points, K, W, H, radius, zmin, zmax, max_depth, nclasses1, frames, d_desc, m_desc, v_desc):
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    _G.update(points=np.asarray(points, dtype=np.float64), K=K, W=W, H=H, radius=radius, zmin=zmin, zmax=zmax,
              max_depth=max_depth, nclasses1=nclasses1, frames=frames)
    _G["keep"] = []
    for key, desc in (("depth", d_desc), ("mask", m_desc), ("votes", v_desc)):
        if isinstance(desc, tuple):
            shm, arr = _attach(desc)
            _G["keep"].append(shm)
            _G[key] = arr
        else:
            _G[key] = desc   # single-process mode: the arrays themselves


def _frame(f):
    g = _G
    quat, t, eye, look, nrm = g["frames"][f]
    idx, pix = orc.fuse_frame_visibility(g["points"], g["K"], g["W"], g["H"], quat, t, eye, look, nrm, g["depth"][f], 0,
                                         g["radius"], g["zmin"], g["zmax"], g["max_depth"])
    cls = g["mask"][f].reshape(-1)[pix]
    return idx.astype(np.int32), cls.astype(np.uint8)


def _segment_rows(ab):
    a, b = ab
    g = _G
    return a, orc.segment(g["votes"][a:b], g["nclasses1"] - 1, 0.5, None)


def _splat(f):
    g = _G
    quat, t, eye, look, nrm = g["frames"][f]
    return f, orc.zero_border(orc.zbuffer_splat(g["points"], g["K"], g["W"], g["H"], quat, t, eye, look, nrm, g["max_depth"]), 10)


def _share(arr):
    shm = shared_memory.SharedMemory(create=True, size=max(arr.nbytes, 1))
    view = np.ndarray(arr.shape, dtype=arr.dtype, buffer=shm.buf)
    view[...] = arr
    return shm, view, (shm.name, arr.shape, arr.dtype.str)


class CpuFusion:
    """Worker processes holding the (sub-sampled) cloud and the sample frames; `run` executes one pass over them.
    `depths=None`: the sample's depth images are rendered first by the oracle's own z-buffer splat of these points
    (`render_depth`), so that the CPU arm needs nothing from the CUDA library."""

    def __init__(self, points, K, W, H, wxyz, t, depths, masks, radius=0.05, zmin=0.1, zmax=4.0, max_depth=4.0,
                 nclasses1=134, workers=None):
        self.cores = len(os.sched_getaffinity(0))
        self.workers = max(1, min(self.cores, len(t), 64) if workers is None else workers)
        self.npoints, self.nframes, self.nclasses1 = len(points), len(t), nclasses1
        eyes, looks, nrms = orc.frustum_data(K, W, H, wxyz, t)
        frames = [(wxyz[f], t[f], eyes[f], looks[f], nrms[f]) for f in range(len(t))]
        if depths is None:
            depths = np.zeros((len(t), H, W), dtype=np.uint16)
        depths, masks = np.ascontiguousarray(depths), np.ascontiguousarray(masks)
        self._shm = []
        if self.workers > 1:
            d_shm, self.depths, d_desc = _share(depths)
            m_shm, self.masks, m_desc = _share(masks)
            v_shm, self.votes, v_desc = _share(np.zeros((self.npoints, nclasses1), dtype=np.int32))
            self._shm = [d_shm, m_shm, v_shm]
            init = (points, K, W, H, radius, zmin, zmax, max_depth, nclasses1, frames, d_desc, m_desc, v_desc)
            self.pool = mp.get_context("spawn").Pool(self.workers, initializer=_init, initargs=init)
        else:
            self.pool = None
            self.depths, self.masks = depths, masks
            self.votes = np.zeros((self.npoints, nclasses1), dtype=np.int32)
            _init(points, K, W, H, radius, zmin, zmax, max_depth, nclasses1, frames, self.depths, self.masks, self.votes)
        self.last = {}

    def render_depth(self):
        """Depth of the sample frames = oracle z-buffer splat of the sample points (uint16 mm, 10-px zero border)."""
        it = self.pool.imap_unordered(_splat, range(self.nframes)) if self.pool else map(_splat, range(self.nframes))
        for f, d in it:
            self.depths[f] = d
        return self.depths

    def run(self):
        """One pass; returns (seconds, votes int32 [N, C1], labels int64 [N]).  `self.last` holds the split:
        seconds of the per-frame stage (project + visibility + vote) and of `segment`."""
        t0 = time.perf_counter()
        votes = self.votes
        votes[...] = 0
        it = self.pool.imap_unordered(_frame, range(self.nframes)) if self.pool else map(_frame, range(self.nframes))
        for idx, cls in it:
            if len(idx):
                votes[idx, cls] += 1                                  # voting.py:98
        t1 = time.perf_counter()
        if self.pool:
            labels = np.empty(self.npoints, dtype=np.int64)
            step = -(-self.npoints // (4 * self.workers))
            for a, lab in self.pool.imap_unordered(_segment_rows, [(a, min(a + step, self.npoints)) for a in range(0, self.npoints, step)]):
                labels[a:a + len(lab)] = lab
        else:
            labels = orc.segment(votes, self.nclasses1 - 1, 0.5, None)
        t2 = time.perf_counter()
        self.last = {"frames_s": t1 - t0, "segment_s": t2 - t1}
        return t2 - t0, votes.copy(), labels

    def close(self):
        if self.pool:
            self.pool.close()
            self.pool.join()
            self.pool = None
        for s in self._shm:
            try:
                s.close()
                s.unlink()
            except FileNotFoundError:
                pass
        self._shm = []


def measure(points, K, W, H, wxyz, t, depths, masks, radius, zmin, zmax, max_depth, nclasses1, warmup=0, steps=1,
            single_frames=8):
    """Both CPU figures on one sample.  Returns (dict for the JSON line, votes, labels of the all-cores pass).
    The single-process figure runs on the first `single_frames` sample frames (it is ~cores x slower)."""
    cpu = CpuFusion(points, K, W, H, wxyz, t, depths, masks, radius, zmin, zmax, max_depth, nclasses1)
    if depths is None:
        depths = cpu.render_depth().copy()
    for _ in range(warmup):
        cpu.run()
    secs, split = [], []
    for _ in range(max(1, steps)):
        s, votes, labels = cpu.run()
        secs.append(s)
        split.append(dict(cpu.last))
    cores, workers = cpu.cores, cpu.workers
    cpu.close()
    sec = float(np.mean(secs))
    nf1 = max(1, min(single_frames, len(t)))
    one = CpuFusion(points, K, W, H, wxyz[:nf1], t[:nf1], depths[:nf1], masks[:nf1], radius, zmin, zmax, max_depth, nclasses1, workers=1)
    s1, _, _ = one.run()
    one.close()
    pv = len(points) * len(t)
    return {
        "value": pv / sec, "seconds": sec, "cores": workers, "host_cores_available": cores,
        "frames_stage_seconds": float(np.mean([d["frames_s"] for d in split])),
        "segment_seconds": float(np.mean([d["segment_s"] for d in split])),
        "single_process": {"value": len(points) * nf1 / s1, "seconds": s1, "frames": nf1, "cores": 1,
                           "note": "as shipped: the reference has no parallelism (timed like get3DSeg.py:75,85)"},
    }, votes, labels, depths
