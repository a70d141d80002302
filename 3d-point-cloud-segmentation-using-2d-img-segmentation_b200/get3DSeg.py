"""Drop-in for the reference's `get3DSeg.py` entry points on the label-fusion path: same function names, arguments,
defaults, return values and on-disk outputs; vote accumulation and label resolve run on the GPU.

Covered (reference lines): `segment` `get3DSeg.py:18-116`, `remove_classes` `:118-221`, `semantic_viz` `:224-286`,
`panoptic_viz` `:289-347`, `load_semantic_segmentation` `:350-355`, `load_csv` `:357-367`, `master_classes`
`:369-475`.  Open3D is neither required nor used: point clouds are written / read by a small binary-PLY codec and
the GUI windows the reference opens are omitted.  Connected components (`split_into_instances`), box membership
and vote / resolve run on the GPU; the rest is file and dictionary bookkeeping.
"""
from __future__ import annotations

import json
import os
import time
from pathlib import Path

import numpy as np

from .Fusion3DSeg.fusion import Fusion
from .Fusion3DSeg.merge_intersecting_bb import _box_corners, fit_obb, merge_bb
from .Fusion3DSeg.segUtils.cv import split_into_instances
from .Fusion3DSeg.segUtils.voting import VotingSegmentation

# the reference looks for these two files next to its own checkout (`get3DSeg.py:376-377`); they are not part of it
CLASSES_CSV = Path(os.path.dirname(__file__)).parent / 'classes.csv'
CLASSES_META = Path(os.path.dirname(__file__)).parent / 'classes_meta.json'


def write_ply(path, points, colors=None, normals=None):
    """Binary little-endian PLY (x y z [nx ny nz] [red green blue]) -- stands in for o3d.io.write_point_cloud."""
    points = np.asarray(points, dtype=np.float64)
    n = len(points)
    fields = [("x", "<f8"), ("y", "<f8"), ("z", "<f8")]
    header = ["ply", "format binary_little_endian 1.0", f"element vertex {n}", "property double x", "property double y",
              "property double z"]
    if normals is not None:
        fields += [("nx", "<f8"), ("ny", "<f8"), ("nz", "<f8")]
        header += ["property double nx", "property double ny", "property double nz"]
    if colors is not None:
        fields += [("red", "u1"), ("green", "u1"), ("blue", "u1")]
        header += ["property uchar red", "property uchar green", "property uchar blue"]
    header.append("end_header")
    rec = np.zeros(n, dtype=fields)
    rec["x"], rec["y"], rec["z"] = points[:, 0], points[:, 1], points[:, 2]
    if normals is not None:
        nn = np.asarray(normals, dtype=np.float64)
        rec["nx"], rec["ny"], rec["nz"] = nn[:, 0], nn[:, 1], nn[:, 2]
    if colors is not None:
        c = np.clip(np.asarray(colors, dtype=np.float64) * 255.0, 0, 255).astype(np.uint8)
        rec["red"], rec["green"], rec["blue"] = c[:, 0], c[:, 1], c[:, 2]
    with open(path, "wb") as fp:
        fp.write(("\n".join(header) + "\n").encode("ascii"))
        fp.write(rec.tobytes())


def read_ply_points(path):
    """xyz of a binary little-endian PLY written by `write_ply` (stands in for o3d.io.read_point_cloud)."""
    with open(path, "rb") as fp:
        props, n = [], 0
        while True:
            line = fp.readline().decode("ascii").strip()
            if line.startswith("element vertex"):
                n = int(line.split()[-1])
            elif line.startswith("property"):
                _, typ, name = line.split()
                props.append((name, {"double": "<f8", "float": "<f4", "uchar": "u1"}[typ]))
            elif line == "end_header":
                break
        rec = np.frombuffer(fp.read(), dtype=props, count=n)
    return np.stack([rec["x"], rec["y"], rec["z"]], axis=1).astype(np.float64)


class _Pcd:
    """Minimal stand-in for the o3d PointCloud handed to `merge_bb` (only `.points` is used there)."""

    def __init__(self, points):
        self.points = points


def semantic_viz(points, classes, nclasses, votes=None, coco_data=None, outdir='./'):
    """Reference `semantic_viz` (`get3DSeg.py:224-286`): dumps classes.npy (+votes.npy), pcd.ply and info.json with
    per-class point counts.  Returns (colors, None, palette, info) -- no Open3D object is built."""
    outdir = Path(outdir)
    outdir.mkdir(exist_ok=True, parents=True)
    if votes is not None:
        np.save(outdir / 'votes.npy', votes)
    np.save(outdir / 'classes.npy', classes)
    if coco_data is not None:
        with open(coco_data, 'r') as fp:
            coco_classes = json.load(fp)['stuff_classes']
    else:
        coco_classes = [str(i) for i in range(nclasses)]
    coco_classes = list(coco_classes) + ['unclassified']
    palette = np.vstack((np.random.uniform(0, 1, size=(nclasses, 3)), np.zeros((1, 3))))
    class_ids, counts = np.unique(classes, return_counts=True)
    colors = np.zeros((len(classes), 3))
    for c in class_ids:
        colors[classes == c] = palette[min(int(c), nclasses)]
    write_ply(outdir / 'pcd.ply', points, colors)

    def tocss(clr):
        return "#" + "".join(hex(int(c)).replace('0x', '').zfill(2) for c in clr)

    info = [{'category_id': int(c), 'name': coco_classes[min(int(c), nclasses)], 'area': int(a),
             'hexcolor': tocss((palette[min(int(c), nclasses)] * 255).astype(int))} for c, a in zip(class_ids, counts)]
    with open(outdir / 'info.json', 'w') as fp:
        json.dump(info, fp, indent=4)
    return colors, None, palette, info


def load_semantic_segmentation(semantic_dir):
    votes = np.load(os.path.join(semantic_dir, 'votes.npy'))
    classes = np.load(os.path.join(semantic_dir, 'classes.npy'))
    with open(os.path.join(semantic_dir, 'info.json'), 'r') as fp:
        info = json.load(fp)
    return votes, classes, classes, np.unique(classes), info


def load_csv(data_path):
    """Reference `load_csv` (`get3DSeg.py:357-367`)."""
    import pandas as pd
    df = pd.read_csv(data_path)
    class_id = df['Class_ID'].tolist()
    flag_removal = np.bool_(df['flag_objremoval'].tolist())
    building_classes = [class_id[i] for i in np.where(flag_removal == False)[0]]  # noqa: E712
    return class_id, df['Parent'].tolist(), df['Parent_ID'].tolist(), df['flag_infojson'].tolist(), building_classes


def _coco_meta():
    p = Path(os.path.dirname(__file__)).parent / 'deeplearning' / 'segmentation' / 'mask2former' / 'coco_meta.json'
    return p if p.is_file() else None


def segment(dirname, mask_dir, threshold=0.5, nclasses=133, filter_classes=[86, 114, 115], min_pts_per_inst=100,
            verbose=True):
    """Semantic segmentation of the fused cloud from 2D masks + uv2pt lookups (reference `get3DSeg.py:18-116`).

    Writes dirname/segmentation/{votes.npy, classes.npy, info.json, pcd.ply}.  Returns `(votes, classes)` when the
    fusion directory has no adjacency list -- exactly the reference's early return (`get3DSeg.py:110`)."""
    dirname = Path(dirname)
    points, norms, colors, nmerges, occurences, nframes, depth_hw, adj = Fusion.load_data(dirname)
    npts = len(points)
    start_time = time.perf_counter()
    voter = VotingSegmentation(npts, depth_hw, mask_dir, dirname / 'fusion' / 'uv2pt', nclasses, votes_file=None)
    votes = voter.vote(resize=True, filename=dirname / 'segmentation' / 'votes.npy', verbose=verbose)
    classes = voter.segment(threshold, filter_classes)
    end_time = time.perf_counter()
    if verbose:
        print(f'Time taken for segmentation = {end_time - start_time} seconds')
    semantic_viz(points, classes, nclasses, votes=None, coco_data=_coco_meta(), outdir=dirname / 'segmentation')
    if adj is None:
        print('No adjacency list available, hence skipping instance seperation.')
        return votes, classes
    insts, ids, pan_info, pan_classes = split_into_instances(classes, adj, nclasses, filter_classes, min_pts_per_inst,
                                                             verbose=verbose)
    panoptic_viz(points, ids, pan_info, dirname / 'panoptic_segmentation', _coco_meta(), colors=None, alpha=1.0)
    master_classes(dirname)


def panoptic_viz(points, ids, idinfo, outdir, coco_data=None, colors=None, alpha=1.0):
    """Reference `panoptic_viz` (`get3DSeg.py:289-347`): dumps ids.npy, info.json (with hexcolor / name) and pcd.ply."""
    outdir = Path(outdir)
    outdir.mkdir(exist_ok=True, parents=True)
    np.save(outdir / 'ids.npy', ids)
    classnames = None
    if coco_data is not None:
        with open(coco_data, 'r') as fp:
            classnames = list(json.load(fp)['stuff_classes']) + ['unclassified']
    allids = np.unique(ids)
    idinfo = [idinfo[i] for i in allids]
    colors = np.zeros((len(points), 3)) if colors is None else np.array(colors, dtype=np.float64)
    palette = np.random.uniform(0, 1, size=(len(allids), 3))
    for id_, info, clr in zip(allids, idinfo, palette):
        info['hexcolor'] = "#" + "".join(hex(int(c)).replace('0x', '').zfill(2) for c in (clr * 255).astype(int))
        info['name'] = classnames[info['category_id']] if classnames is not None else str(info['category_id'])
        m = ids == id_
        colors[m] = (1 - alpha) * colors[m] + alpha * clr
    with open(outdir / 'info.json', 'w') as fp:
        json.dump(idinfo, fp, indent=4)
    write_ply(outdir / 'pcd.ply', points, colors)
    return colors, None, palette, idinfo


def master_classes(dirname):
    """Reference `master_classes` (`get3DSeg.py:369-475`): attaches parent ids / names / box corners from `classes.csv`
    + `classes_meta.json`, rewrites both info.json files and segmentation/final_pcd.ply, then merges intersecting
    instance boxes (`merge_bb`, `:475`).  Oriented boxes come from `fit_obb` (batched GPU fit on the stated box model, see
    `oracle.fit_box`; Open3D's hull-based fit is unpinned); everything else is pinned against the files the unmodified reference
    writes for the same inputs (tests/golden/make_golden_master.py)."""
    import torch
    dirname = Path(dirname)
    class_id, parent_name, parent_id, flag_infojson, _ = load_csv(CLASSES_CSV)
    points = read_ply_points(dirname / 'panoptic_segmentation' / 'pcd.ply')
    ids = np.load(dirname / 'panoptic_segmentation' / 'ids.npy')
    classes = np.load(dirname / 'segmentation' / 'classes.npy')
    parent_classes = classes.copy()
    with open(dirname / 'panoptic_segmentation' / 'info.json', 'r') as fp:
        info_pan = json.load(fp)
    with open(dirname / 'segmentation' / 'info.json', 'r') as fp:
        info_sem = json.load(fp)
    with open(CLASSES_META, 'r') as fp:
        classes_meta = json.load(fp)
    palette = np.divide(np.array(classes_meta['colors']), 255)

    def tocss(clr):
        return "#" + "".join(hex(int(c)).replace('0x', '').zfill(2) for c in clr)

    pts_dev = torch.as_tensor(points).cuda()
    ids_dev = torch.as_tensor(ids).cuda()
    final_info, area_unclassified, unclassified_instance = [], 0, None
    for info in info_pan:                                                          # get3DSeg.py:422-452
        mask = ids == info['id']
        if info['category_id'] in class_id:
            k = class_id.index(info['category_id'])
            info['parent_id'], info['parent_name'] = parent_id[k], parent_name[k]
            info['parent_hexcolor'] = tocss((palette[info['parent_id']] * 255).astype(int))
            if info['category_id'] == 133:
                unclassified_instance = info['id']
                info['bbox'] = None
            else:
                info['bbox'] = _box_corners(fit_obb(pts_dev[ids_dev == info['id']]).cpu().numpy())
            if flag_infojson[k]:
                final_info.append(info)
        else:
            area_unclassified += int(np.count_nonzero(mask))
            info['parent_id'] = info['parent_name'] = info['parent_hexcolor'] = info['bbox'] = None
    if unclassified_instance is not None and unclassified_instance < len(final_info):
        final_info[unclassified_instance]['area'] += area_unclassified               # index-as-id, as :450
    for info in info_sem:                                                          # :454-462
        mask = classes == info['category_id']
        if info['category_id'] in class_id:
            k = class_id.index(info['category_id'])
            info['parent_id'], info['parent_name'] = parent_id[k], parent_name[k]
            info['parent_hexcolor'] = tocss((palette[info['parent_id']] * 255).astype(int))
            parent_classes[mask] = int(info['parent_id'])
        else:
            parent_classes[mask] = classes_meta['classes'].index('unclassified')
    colors = np.zeros_like(points)
    for c in np.unique(parent_classes):
        colors[parent_classes == c] = palette[c]
    write_ply(dirname / 'segmentation' / 'final_pcd.ply', points, colors)
    with open(dirname / 'segmentation' / 'info.json', 'w') as fp:
        json.dump(info_sem, fp, indent=4)
    with open(dirname / 'panoptic_segmentation' / 'info.json', 'w') as fp:
        json.dump(info_pan, fp, indent=4)
    merge_bb(dirname, final_info, ids, _Pcd(points))                               # :475


def remove_classes(dirname, mask_dir, keep_classes, threshold=0.75, nclasses=133, verbose=True):
    """Mask of the points to keep (reference `get3DSeg.py:118-221`).  Like the reference, a `classes.csv` next to the
    package overrides `keep_classes` (`:143-144`); votes are re-used from segmentation/votes.npy when present
    (`:158-164`, including the nclasses = 134 quirk of `voting.py:40`)."""
    if Path(CLASSES_CSV).is_file():
        _, _, _, _, keep_classes = load_csv(CLASSES_CSV)
    dirname = Path(dirname)
    points, norms, colors, nmerges, occurences, nframes, depth_hw, adj = Fusion.load_data(dirname)
    colors = np.zeros((len(points), 3)) if colors is None else np.array(colors, dtype=np.float64)
    colors_org = colors.copy()
    npts = len(points)
    start_time = time.perf_counter()
    votes_file = dirname / 'segmentation' / 'votes.npy'
    votes_file = votes_file if votes_file.is_file() else None
    voter = VotingSegmentation(npts, depth_hw, mask_dir, dirname / 'fusion' / 'uv2pt', nclasses, votes_file=votes_file)
    if votes_file is None:
        voter.vote(resize=True, filename=dirname / 'segmentation' / 'votes.npy', verbose=verbose)
    classes = voter.segment(threshold, None)
    end_time = time.perf_counter()
    if verbose:
        print(f'Time taken for segmentation = {end_time - start_time} seconds')

    removed = np.setdiff1d(np.arange(nclasses), keep_classes)
    removed = np.append(removed, [133, 134])                         # get3DSeg.py:174-175
    remaining_mask = ~np.isin(classes, removed)
    (dirname / 'segmentation').mkdir(exist_ok=True, parents=True)
    np.save(dirname / 'segmentation' / 'remaining_mask.npy', remaining_mask)
    colors[remaining_mask] = [1, 0, 0]
    colors[~remaining_mask] = [0, 0, 1]
    write_ply(dirname / 'segmentation' / 'remaining.ply', points, colors)
    write_ply(dirname / 'segmentation' / 'cleaned.ply', np.asarray(points)[remaining_mask], colors_org[remaining_mask],
              None if norms is None else np.asarray(norms)[remaining_mask])
    removed_point_classes = classes.copy()
    removed_point_classes[remaining_mask] = 133
    removed_point_classes[removed_point_classes == 134] = 133
    semantic_viz(points, removed_point_classes, nclasses, votes=None, coco_data=_coco_meta(),
                 outdir=dirname / 'segmentation' / 'removed_objects_info')
    return remaining_mask
