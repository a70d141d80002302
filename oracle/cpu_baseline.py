"""CPU arm of the benchmark -- TEST / MEASUREMENT INFRASTRUCTURE, NOT PRODUCT CODE.

Times the numpy restatement of the reference's label-fusion path (`oracle/f3d_oracle.py`, kind = "port": the
reference itself is Python + third-party modules that are absent from this image and `/root/reference` does not
exist on the GPU box) on the host cores.  The reference has no parallelism of its own; as SURVEY 8(d) prescribes
the frames are sharded over worker processes (`multiprocessing`, spawn) and the per-frame (point, class) hits
are accumulated with `votes[idx, cls] += 1` (`segUtils/voting.py:98`) in the parent.  Only `bench.py`
(cpu_baseline leg and `--impl reference`) imports this module.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np

from . import f3d_oracle as orc

_G = {}


def _init(points, K, W, H, radius, zmin, zmax, max_depth):
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    _G.update(points=np.asarray(points, dtype=np.float64), K=K, W=W, H=H, radius=radius, zmin=zmin, zmax=zmax,
              max_depth=max_depth)


def _frame(args):
    quat, t, eye, look, nrm, depth, mask = args
    g = _G
    idx, pix = orc.fuse_frame_visibility(g["points"], g["K"], g["W"], g["H"], quat, t, eye, look, nrm, depth, 0,
                                         g["radius"], g["zmin"], g["zmax"], g["max_depth"])
    cls = mask.reshape(-1)[pix]
    return idx.astype(np.int32), cls.astype(np.uint8)


class CpuFusion:
    """Pool of worker processes holding the (sub-sampled) cloud; `run` executes one pass over the sample frames."""

    def __init__(self, points, K, W, H, wxyz, t, depths, masks, radius=0.05, zmin=0.1, zmax=4.0, max_depth=4.0,
                 nclasses1=134, workers=None):
        self.cores = len(os.sched_getaffinity(0))
        self.workers = max(1, min(self.cores, len(t), 64) if workers is None else workers)
        self.npoints, self.nframes, self.nclasses1 = len(points), len(t), nclasses1
        eyes, looks, nrms = orc.frustum_data(K, W, H, wxyz, t)
        self.tasks = [(wxyz[f], t[f], eyes[f], looks[f], nrms[f], depths[f], masks[f]) for f in range(len(t))]
        init = (points, K, W, H, radius, zmin, zmax, max_depth)
        if self.workers > 1:
            self.pool = mp.get_context("spawn").Pool(self.workers, initializer=_init, initargs=init)
        else:
            self.pool = None
            _init(*init)

    def run(self):
        """One pass; returns (seconds, votes int32 [N, C1], labels int64 [N])."""
        t0 = time.perf_counter()
        votes = np.zeros((self.npoints, self.nclasses1), dtype=np.int32)
        it = self.pool.imap_unordered(_frame, self.tasks) if self.pool else map(_frame, self.tasks)
        for idx, cls in it:
            if len(idx):
                votes[idx, cls] += 1
        labels = orc.segment(votes, self.nclasses1 - 1, 0.5, None)
        return time.perf_counter() - t0, votes, labels

    def close(self):
        if self.pool:
            self.pool.close()
            self.pool.join()
            self.pool = None
