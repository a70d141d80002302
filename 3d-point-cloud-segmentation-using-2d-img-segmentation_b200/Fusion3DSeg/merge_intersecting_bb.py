"""Drop-in for `Fusion3DSeg/merge_intersecting_bb.py`: instance-box intersection and merging on the GPU.

Two contracts (SURVEY a-14 / a-15):

* `merge_boxes(lo, hi, group, area)` -- the batched kernel path of the north star (config C5): closed-interval
  AABB overlap of `check_intersection` (reference `:44-56`, predicate `:51-53`, same-category gate `:49`) over
  all pairs, closed transitively with a GPU union-find; label = smallest box index of the component.
* `merge_bb(dir_name, info_sem, id_info_per_point, pcd)` -- the reference's sequential, order-dependent driver
  (`:103-137`) with all its quirks (loop index used as instance id `:70,113`; shrinking-list guards `:79`; early
  `return` on a < 4 point instance `:83-84`; `del` without index correction `:118-120`), emulated on the host
  while every geometric predicate (oriented-box membership of the whole cloud, `:75-76,86-87`) runs on the GPU.
  Open3D's `OrientedBoundingBox.create_from_points` is not available in this image; boxes are fitted by
  `fit_obb` (PCA axes + extents, the covariance variant of Open3D's algorithm) -- "parity unpinned" for the fit,
  the membership rule |(p-c).axis_k| <= extent_k/2 is Open3D's.
"""
from __future__ import annotations

import json
import time
from pathlib import Path

import numpy as np
import torch

from .. import engine
from .._lib import require_cuda


def merge_boxes(lo, hi, group, area=None):
    """lo, hi [B,3] float64, group [B] int -> (labels int32 [B], edges int32 [E,2], merged_area or None).
    labels[i] = smallest index of i's connected component under the same-group closed AABB overlap."""
    edges = engine.box_pairs_aabb(lo, hi, group)
    labels = engine.union_find(len(lo), edges)
    merged = None
    if area is not None:
        a = engine.as_cuda(area, torch.int64)
        merged = torch.zeros_like(a).index_add_(0, labels.to(torch.int64), a)
    return labels, edges, merged


def check_intersection(lo, hi, group):
    """All intersecting same-group box pairs (i < j), sorted -- `check_intersection` (`:44-56`) for every id1 at once."""
    e = engine.box_pairs_aabb(lo, hi, group).cpu().numpy().astype(np.int64)
    return np.unique(e, axis=0) if len(e) else e.reshape(0, 2)


def fit_obb(points_dev: torch.Tensor):
    """Oriented box of a point set on the GPU: centre, rotation (columns = axes, sorted by decreasing variance,
    right-handed) and extents.  Returns a float64 [15] device tensor (centre, R row-major, extent)."""
    p = points_dev.to(torch.float64)
    mean = p.mean(0)
    q = p - mean
    cov = (q.T @ q) / max(len(p) - 1, 1)
    evals, evecs = torch.linalg.eigh(cov)
    R = evecs[:, [2, 1, 0]].clone()
    R[:, 2] = torch.linalg.cross(R[:, 0], R[:, 1])
    proj = q @ R
    mn, mx = proj.min(0).values, proj.max(0).values
    centre = mean + R @ ((mn + mx) * 0.5)
    return torch.cat([centre, R.reshape(-1), mx - mn])


def _box_corners(box15: np.ndarray):
    c, R, e = box15[:3], box15[3:12].reshape(3, 3), box15[12:]
    s = np.array([[sx, sy, sz] for sx in (-0.5, 0.5) for sy in (-0.5, 0.5) for sz in (-0.5, 0.5)])
    return (c[None, :] + (s * e[None, :]) @ R.T).tolist()


def merge_bb(dir_name, info_sem, id_info_per_point, pcd):
    """Same call and side effects as the reference `merge_bb` (`merge_intersecting_bb.py:103-137`): mutates
    `info_sem` and `id_info_per_point`, writes panoptic_segmentation/{final_info.json, ids.npy}."""
    dev = require_cuda()
    len_info_sem = len(info_sem)
    pts_np = np.ascontiguousarray(np.asarray(pcd.points if hasattr(pcd, "points") else pcd, dtype=np.float64))
    pts = torch.as_tensor(pts_np).to(dev)
    ids = torch.as_tensor(np.ascontiguousarray(id_info_per_point)).to(dev)
    start_time = time.perf_counter()

    def inside_of(instance_id):
        """None if the instance has < 4 points (`:72,83`), else the uint8 [N] membership of the whole cloud (`:75-76`)."""
        sel = ids == instance_id
        if int(sel.sum()) < 4:
            return None
        return engine.obb_contains(pts, fit_obb(pts[sel])[None, :])[0]

    L = len(info_sem)
    for id1 in range(1, L):                                               # :113
        hits = []
        a = inside_of(id1)
        if a is not None:
            for id2 in range(1, L):                                       # :78
                if id1 != id2 and id2 < len(info_sem) - 1 and id1 < len(info_sem) - 1:   # :79
                    if info_sem[id1]["parent_id"] == info_sem[id2]["parent_id"]:        # :80
                        b = inside_of(id2)
                        if b is None:
                            break                                         # early return, :83-84
                        if bool((a & b).any()):                           # :88-90
                            hits.append(id2)
        if hits:
            for hb in hits:                                               # update_id_info, :58-62
                info_sem[id1]["area"] += info_sem[hb]["area"]
                ids[ids == hb] = id1
            for i in hits:                                                # :118-120
                if i < len(info_sem):
                    del info_sem[i]

    for k in range(1, len(info_sem)):                                     # :122-128
        sel = ids == info_sem[k]["id"]
        if int(sel.sum()) > 4:
            info_sem[k]["bbox"] = _box_corners(fit_obb(pts[sel]).cpu().numpy())

    id_info_per_point[...] = ids.cpu().numpy()
    end_time = time.perf_counter()
    print(f'Time taken for merging {len_info_sem} to {len(info_sem)} Bounding boxes = {end_time - start_time} seconds')
    out = Path(dir_name) / "panoptic_segmentation"
    out.mkdir(exist_ok=True, parents=True)
    with open(out / "final_info.json", 'w') as fp:
        json.dump(info_sem, fp, indent=4)
    with open(out / "ids.npy", 'wb') as fi:
        np.save(fi, id_info_per_point)
