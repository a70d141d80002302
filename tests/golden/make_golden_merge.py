"""Golden vectors for the instance-box merge (SURVEY a-14 / a-15), produced by the UNMODIFIED reference module
`/root/reference/Fusion3DSeg/merge_intersecting_bb.py` in the build container.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_merge.py

The module's two third-party imports that are absent from this image are replaced by `oracle/refshim/open3d` (the three
`OrientedBoundingBox` calls, backed by a STATED box model: "pca" and "aabb", see that file) and
`oracle/refshim/skspatial` (`Line.project_point`).  What these vectors pin is the reference's driver logic --
`merge_bb` with its index-as-id / shrinking-list / early-return quirks, `update_id_info`, `check_intersection_open3d`,
`cal_min_max`, `check_intersection` -- run as shipped; Open3D's hull-based box FIT stays unpinned (stated in DESIGN.md).
Output: tests/golden/g7_merge.json.
"""
import contextlib
import copy
import io
import json
import os
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
REF = Path(os.environ.get("F3D_REFERENCE", "/root/reference"))
sys.dont_write_bytecode = True
sys.path.insert(0, str(ROOT / "oracle" / "refshim"))
sys.path.insert(0, str(REF))

import open3d as o3d  # noqa: E402  (the shim)
from Fusion3DSeg import merge_intersecting_bb as ref  # noqa: E402


def scenario(seed, kind):
    """Instances 1..8 along x (id 0 = background), ~60 points each in rotated boxes.
    kind "chain": overlap chain 1-2-3, overlapping pair 5-6, instance 7 has 3 points, instance 4 isolated (SURVEY a-14);
    kind "parents": same geometry, but instances 2 and 6 carry another parent_id (the `:80` gate);
    kind "dense": every instance overlaps its neighbour (long cascade of deletions)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    centres = {1: 0.0, 2: 0.9, 3: 1.8, 4: 4.0, 5: 6.0, 6: 6.8, 7: 9.0, 8: 11.0}
    if kind == "dense":
        centres = {k: 0.85 * (k - 1) for k in range(1, 9)}
    pts, ids = [], []
    for k, cx in centres.items():
        n = 3 if (k == 7 and kind != "dense") else 60
        yaw = rng.uniform(-0.3, 0.3)
        R = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]])
        local = rng.uniform(-0.5, 0.5, (n, 3)) * np.array([1.1, 0.6, 0.4])
        pts.append(local @ R.T + np.array([cx, 0.3 * np.sin(k), 0.5]))
        ids += [k] * n
    bg = rng.uniform(-1.0, 1.0, (200, 3)) * np.array([8.0, 3.0, 0.2]) + np.array([5.5, 0.0, 3.0])   # ceiling, outside every box
    pts.append(bg)
    ids += [0] * len(bg)
    pts = np.round(np.concatenate(pts), 6) + 0.123456      # no exact zeros (cal_min_max drops zero coordinates, :23-25)
    ids = np.asarray(ids, dtype=np.int64)
    info = [{"id": k, "category_id": 100 + (k % 2), "parent_id": 7, "parent_name": "furniture", "area": 10 * k + 3, "isthing": True}
            for k in range(0, 9)]
    if kind == "parents":
        info[2]["parent_id"] = 9
        info[6]["parent_id"] = 9
    return pts, ids, info


def run_merge(pts, ids, info, model):
    o3d.BOX_MODEL = model
    info, ids = copy.deepcopy(info), ids.copy()
    pcd = o3d.geometry.PointCloud(pts)
    with tempfile.TemporaryDirectory() as td:
        (Path(td) / "panoptic_segmentation").mkdir()
        with contextlib.redirect_stdout(io.StringIO()):
            ref.merge_bb(Path(td), info, ids, pcd)                    # mutates info / ids, writes the two files
        saved_ids = np.load(Path(td) / "panoptic_segmentation" / "ids.npy")
        saved_info = json.loads((Path(td) / "panoptic_segmentation" / "final_info.json").read_text())
    assert np.array_equal(saved_ids, ids) and saved_info == json.loads(json.dumps(info))
    return info, ids


def main():
    out = {"note": "reference merge_intersecting_bb.py run unmodified on oracle/refshim box models", "cases": []}
    for seed, kind in ((11, "chain"), (12, "parents"), (13, "dense"), (14, "chain")):
        pts, ids, info = scenario(seed, kind)
        case = {"seed": seed, "kind": kind, "points": pts.tolist(), "ids": ids.tolist(), "info_sem": info, "models": {}}
        for model in ("pca", "aabb"):
            fin_info, fin_ids = run_merge(pts, ids, info, model)
            o3d.BOX_MODEL = model
            # cal_min_max (:15-42) and check_intersection (:44-56, incl. its reset-inside-the-loop behaviour) as shipped
            id_list = [d["id"] for d in info]
            mm = {str(k): [np.asarray(v).tolist() for v in ref.cal_min_max(k, ids, pts)] for k in id_list if (ids == k).sum() >= 4}
            ci = {}
            for id1 in range(1, len(id_list)):
                if (ids == id_list[id1]).sum() >= 4 and all((ids == id_list[j]).sum() >= 4 or info[id1]["category_id"] != info[j]["category_id"]
                                                             for j in range(1, len(id_list)) if j != id1):
                    ci[str(id1)] = ref.check_intersection(id1, id_list, ids, pts, info)
            # check_intersection_open3d (:68-91) called directly on the INITIAL state, every id1 (incl. the < 4 point instance
            # and the early `return` an id2 with < 4 points causes)
            pcd0 = o3d.geometry.PointCloud(pts)
            cio = {str(id1): [int(v) for v in ref.check_intersection_open3d(id1, id_list, ids, pts, pcd0, info)]
                   for id1 in range(1, len(id_list))}
            case["models"][model] = {"final_info": fin_info, "final_ids": fin_ids.tolist(), "cal_min_max": mm, "check_intersection": ci,
                                     "check_intersection_open3d": cio}
        out["cases"].append(case)
        print(kind, seed, {m: [d["id"] for d in case["models"][m]["final_info"]] for m in case["models"]},
              {m: [d["area"] for d in case["models"][m]["final_info"]] for m in case["models"]})
    (HERE / "g7_merge.json").write_text(json.dumps(out))
    print("wrote", HERE / "g7_merge.json", (HERE / "g7_merge.json").stat().st_size, "bytes")


if __name__ == "__main__":
    main()
