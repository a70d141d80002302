"""Drop-in for the reference's `get3DSeg.py` entry points on the label-fusion path: same function names, arguments,
defaults, return values and on-disk outputs; vote accumulation and label resolve run on the GPU.

Covered (reference lines): `segment` `get3DSeg.py:18-116` (through the semantic dumps; see below),
`remove_classes` `:118-221`, `semantic_viz` `:224-286`, `load_semantic_segmentation` `:350-355`, `load_csv`
`:357-367`.  Open3D is neither required nor used: point clouds are written by a small binary-PLY writer and
the GUI windows the reference opens when `verbose` are omitted.  The instance split (`split_into_instances`,
SURVEY 8(f) rank 1) is not on the GPU yet: when the fusion directory carries an adjacency list `segment`
raises NotImplementedError after writing the semantic outputs, instead of silently running a CPU version.
"""
from __future__ import annotations

import json
import os
import time
from pathlib import Path

import numpy as np

from .Fusion3DSeg.fusion import Fusion
from .Fusion3DSeg.segUtils.voting import VotingSegmentation


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
    raise NotImplementedError(
        "instance split (split_into_instances, cv.py:402-500) and the panoptic dumps are the next rows of the hot-path "
        "table and are not built yet; remove fusion/adj.pkl to get the semantic (votes, classes) result")


def remove_classes(dirname, mask_dir, keep_classes, threshold=0.75, nclasses=133, verbose=True):
    """Mask of the points to keep (reference `get3DSeg.py:118-221`).  Like the reference, a `classes.csv` next to the
    package overrides `keep_classes` (`:143-144`); votes are re-used from segmentation/votes.npy when present
    (`:158-164`, including the nclasses = 134 quirk of `voting.py:40`)."""
    classes_csv = Path(os.path.dirname(__file__)).parent / 'classes.csv'
    if classes_csv.is_file():
        _, _, _, _, keep_classes = load_csv(classes_csv)
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
