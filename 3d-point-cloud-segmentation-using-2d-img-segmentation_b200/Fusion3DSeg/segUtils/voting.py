"""Drop-in for `Fusion3DSeg/segUtils/voting.py` of the reference: same class, arguments, attributes and return
types; mask gather + vote + label resolve run on the GPU (f3d_vote_uv2pt / f3d_resolve_labels)."""
from __future__ import annotations

from pathlib import Path

import cv2
import numpy as np
import torch

from ... import engine
from ..._lib import F3dError, require_cuda


def _votes_to_device(votes) -> torch.Tensor:
    """float64 / int vote matrix (integer valued, as the reference only ever produces) -> int32 device tensor."""
    a = np.asarray(votes)
    if a.ndim != 2:
        raise ValueError("votes must be [npts, nclasses+1]")
    if a.dtype.kind == "f":
        if a.size and (not np.all(a == np.floor(a)) or a.min() < 0 or a.max() >= 2 ** 31):
            raise F3dError("votes must be non-negative integer counts (the GPU path has no float-vote fallback)")
    return torch.as_tensor(np.ascontiguousarray(a.astype(np.int32))).to(require_cuda())


class VotingSegmentation:
    """Voting based 3D point cloud segmentation given 2D masks & uv2pt lookups
    (reference: `Fusion3DSeg/segUtils/voting.py:11-137`)."""

    FRAME_BATCH = 64   # frames staged per host->device copy

    def __init__(self, npts, depth_hw, maskdir, uv2ptdir, nclasses, votes_file=None):
        self._votes_np = None
        if votes_file is None:
            self.npts = npts
            self.depth_hw = depth_hw
            self.nclasses = nclasses
            self._votes_dev = torch.zeros((npts, nclasses + 1), dtype=torch.int32, device=require_cuda())
            self.mask_files, self.uv2pt_files = self._get_filenames(maskdir, uv2ptdir)
            self.nframes = len(self.mask_files)
        else:
            v = np.load(votes_file)
            self._votes_dev = _votes_to_device(v)
            self._votes_np = v
            self.nclasses = v.shape[1]          # reference quirk kept: nclasses+1 when loaded (voting.py:40)

    # -- `votes` keeps the reference's type (float64 [npts, nclasses+1]); the device int32 tensor is the source
    @property
    def votes(self):
        if self._votes_np is None:
            self._votes_np = self._votes_dev.to(torch.float64).cpu().numpy()
        return self._votes_np

    @votes.setter
    def votes(self, value):
        self._votes_dev = _votes_to_device(value)
        self._votes_np = np.asarray(value)

    def _get_filenames(self, maskdir, uv2ptdir):
        """Pair mask and uv2pt files by stem (reference `voting.py:42-55`); sorted for a reproducible order."""
        maskdir, uv2ptdir = Path(maskdir), Path(uv2ptdir)
        mask_names = {p.stem: p for p in maskdir.iterdir() if p.is_file()}
        uv2pt_names = {p.stem: p for p in uv2ptdir.iterdir() if p.is_file()}
        names = sorted(set(mask_names) & set(uv2pt_names))
        return [mask_names[n] for n in names], [uv2pt_names[n] for n in names]

    def _read_data(self, idx):
        mask = cv2.imread(str(self.mask_files[idx]), 0)          # voting.py:66
        uv2pt = np.load(self.uv2pt_files[idx])                   # voting.py:67
        return mask, uv2pt

    def zero(self):
        self._votes_dev.zero_()
        self._votes_np = None

    def vote(self, resize=True, verbose=False, filename=None):
        """Accumulate votes of every frame (reference `voting.py:75-104`).  Returns float64 [npts, nclasses+1]."""
        h, w = self.depth_hw
        dev = self._votes_dev.device
        if verbose:
            print('voting ... ')
        packed = torch.zeros_like(self._votes_dev)
        tag = 1
        for b0 in range(0, self.nframes, self.FRAME_BATCH):
            idxs = range(b0, min(b0 + self.FRAME_BATCH, self.nframes))
            data = [self._read_data(i) for i in idxs]
            if verbose:
                print(f'frame/total = {idxs[-1] + 1}/{self.nframes}, progress = {((idxs[-1] + 1) * 100 / self.nframes):.3}%')
            uv = torch.as_tensor(np.stack([u.astype(np.int32, copy=False).reshape(-1) for _, u in data])).to(dev)
            shapes = {m.shape for m, _ in data}
            groups = [list(range(len(data)))] if len(shapes) == 1 else [[k] for k in range(len(data))]
            masks = torch.empty((len(data), h * w), dtype=torch.uint8, device=dev)
            for g in groups:
                m = torch.as_tensor(np.stack([data[k][0] for k in g])).to(dev)
                if resize and tuple(m.shape[1:]) != (h, w):
                    m = engine.resize_nearest(m, h, w)               # voting.py:93
                masks[g] = m.reshape(len(g), -1)
            if tag + len(data) - 1 > 65535:                          # 16-bit frame tags exhausted: fold and restart
                engine.vote_finalize(packed)
                self._votes_dev += packed
                packed.zero_()
                tag = 1
            engine.vote_uv2pt(packed, uv, masks, tag)                # voting.py:95-98
            tag += len(data)
        engine.vote_finalize(packed)
        self._votes_dev += packed
        self._votes_np = None
        votes = self.votes
        if filename is not None:
            Path(filename).parent.mkdir(exist_ok=True, parents=True)
            np.save(filename, votes)
        return votes

    def segment(self, threshold=0.5, filter_classes=None, votes=None):
        """Classify points given votes (reference `voting.py:106-137`).  Returns int64 [npts]."""
        dev_votes = self._votes_dev if votes is None else _votes_to_device(votes)
        labels = engine.resolve_labels(dev_votes, self.nclasses, threshold, filter_classes)
        return labels.cpu().numpy()
