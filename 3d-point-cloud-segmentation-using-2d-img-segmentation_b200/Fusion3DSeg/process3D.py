"""Drop-in for `Fusion3DSeg/process3D.py` of the reference for a GIVEN cloud.

`process3DSeg` keeps the reference's call (`process3D.py:14-19`), reads the same RTAB cache
(`PointcloudMergeResults/{tofsegment,rtscameradata}_*.pkl` + the per-frame pickles, `process3D.py:23-31`, `fusion.py:17-47`),
writes the same hand-off files (`fusion/uv2pt/<frame>.npy` `fusion.py:326-327`, `fusion/fusion_data.pkl` `fusion.py:360-368`,
`fusion/adj.pkl` `fusion.py:369-377`) and returns `Fusion.load_data`'s 8 values (`process3D.py:63-68`).

What differs, deliberately: the reference BUILDS the cloud here (`Fusion.fuse`, running-mean merging + randomised patch
down-sampling, `fusion.py:134-324`) -- sequential, order dependent and seeded by `np.random.shuffle`, outside the bit-exact
contract and not rebuilt.  This entry associates a cloud that already exists (`cloud=` or an earlier
`fusion_data.pkl` in `output_path`) with every frame on the GPU (cull `fusion.py:254-260` -> project `:266` -> single-pixel
`criterion` distance test `:223-225`), i.e. `stride` = 1 and no normal-cosine term (`:226-227`): both arguments are accepted
for signature compatibility and must be left at values that select this behaviour or are reported.
"""
from __future__ import annotations

import os
import time
from pathlib import Path

import numpy as np
import torch

from .. import engine
from ..fused import FusedLabeler
from .fusion import FrameData, Fusion, parse_rts


def process3DSeg(input_data_path, output_path, radius=0.05, angle=10, stride=10, point_range=(0.1, 4), decimation=1, min_occ=3,
                 verbose=False, cloud=None, chunk=32):
    """Reference `process3DSeg` (`process3D.py:14-68`) on a fixed cloud.  `cloud`: points [N,3] or a dict with 'points'
    and optionally 'normals' / 'colors'; default: the cloud of `output_path/fusion/fusion_data.pkl`."""
    mergeresults_path = os.path.join(input_data_path, 'PointcloudMergeResults')
    if not os.path.exists(mergeresults_path):
        raise FileNotFoundError('tofcameradata not found')                        # process3D.py:28 prints and then fails on `tof`
    ss = [f for f in os.listdir(mergeresults_path) if f.__contains__('tofsegment')][0][:-4]
    subfilename = ss.split('_', 1)[1]
    tof = os.path.join(mergeresults_path, f"tofsegment_{subfilename}.pkl")
    rts = os.path.join(mergeresults_path, f"rtscameradata_{subfilename}.pkl")

    dirname = Path(output_path)
    if cloud is None:
        if not (dirname / 'fusion' / 'fusion_data.pkl').is_file():
            raise NotImplementedError("process3DSeg: pass cloud= (cloud construction, Fusion.fuse fusion.py:134-324, is not rebuilt)")
        pts, norms, clrs, *_ = Fusion.load_data(dirname)
    elif isinstance(cloud, dict):
        pts, norms, clrs = cloud['points'], cloud.get('normals'), cloud.get('colors')
    else:
        pts, norms, clrs = cloud, None, None
    pts = np.ascontiguousarray(np.asarray(pts, dtype=np.float64))

    start = time.perf_counter()
    K, w, h, wxyzs, translations = parse_rts(rts)
    frames = FrameData(tof, point_range, decimation, (h, w))
    nframes = len(frames)
    fl = FusedLabeler(pts, K, w, h, wxyzs[:nframes], translations[:nframes], point_range, radius, max_depth=point_range[1])
    out_dir = dirname / 'fusion' / 'uv2pt'
    out_dir.mkdir(exist_ok=True, parents=True)
    nmerges = torch.zeros(len(pts), dtype=torch.int64, device=fl.points4.device)
    occurences = torch.zeros(len(pts), dtype=torch.int64, device=fl.points4.device)
    for a in range(0, nframes, chunk):
        b = min(a + chunk, nframes)
        names, depths = zip(*[frames.depth_mm(i) for i in range(a, b)])
        uv = fl.uv2pt(torch.as_tensor(np.stack(depths)).to(fl.points4.device), frame_begin=a, frame_end=b)   # int32 [b-a, h*w], -1 = none
        for k in range(b - a):
            hit = uv[k][uv[k] >= 0].to(torch.int64)
            nmerges += torch.bincount(hit, minlength=len(pts))                    # matched pixels per point (`x_merges += matches`, :293)
            occurences[torch.unique(hit)] += 1                                    # frames that saw the point (`x_occ += 1`, :294)
        uv = uv.cpu().numpy()
        for k, name in enumerate(names):
            np.save(out_dir / f'{name}.npy', uv[k])                               # fusion.py:326-327
    end = time.perf_counter()
    nmerges, occurences = nmerges.cpu().numpy(), occurences.cpu().numpy().astype(np.uint32)
    if verbose:
        print(f'\ntotal {h * w * nframes} points from {nframes} frames are associated with {len(pts)} points')
        print(f'time taken for fusion = {(end - start) / 60} minutes')
    if min_occ is not None and verbose:                                           # process3D.py:50-55: computed, printed, discarded
        mask = nmerges >= min_occ
        print(f'remaining points after frame occurence thresholding with {min_occ} = {mask.sum()}')
    if fl.points_rounded and verbose:
        print('note: cloud coordinates were rounded to float32 for the GPU path')
    Fusion.dump_data(dirname, pts, norms, clrs, nmerges, occurences, nframes, (h, w), compute_adjacency=True, ds_radius=radius)
    return Fusion.load_data(dirname)
