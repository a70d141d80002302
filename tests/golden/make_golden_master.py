"""Golden vectors for `get3DSeg.master_classes` (SURVEY a-16), produced by the UNMODIFIED reference `/root/reference/get3DSeg.py`.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_master.py

`master_classes` looks for `classes.csv` / `classes_meta.json` in the PARENT directory of its own file (`get3DSeg.py:376-377`); they are
not part of the reference checkout.  The script therefore loads the reference module from a throw-away copy under /tmp (never
into this repo) next to a small synthetic class table, with the `oracle/refshim` stand-ins for open3d (stated box model "pca", PLY
I/O, no-op window) and skspatial.  Inputs and every file the call writes are stored in tests/golden/g9_master.json.
"""
import contextlib
import importlib.util
import io
import json
import os
import shutil
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
REF = Path(os.environ.get("F3D_REFERENCE", "/root/reference"))
sys.dont_write_bytecode = True

CLASS_TABLE = [  # Class_ID, Parent, Parent_ID, flag_infojson, flag_objremoval
    (133, "unclassified", 5, 1, 0), (10, "furniture", 0, 1, 1), (11, "furniture", 0, 1, 1), (20, "wall", 1, 1, 0), (21, "appliance", 2, 1, 1),
    (30, "floor", 3, 0, 0)]
PALETTE = [[255, 0, 0], [0, 255, 0], [0, 0, 255], [128, 128, 0], [10, 20, 30], [0, 0, 0]]
META_CLASSES = ["furniture", "wall", "appliance", "floor", "misc", "unclassified"]


def write_tables(d):
    (Path(d) / "classes.csv").write_text("Class_ID,Parent,Parent_ID,flag_infojson,flag_objremoval\n" +
                                         "\n".join(",".join(str(v) for v in row) for row in CLASS_TABLE) + "\n")
    (Path(d) / "classes_meta.json").write_text(json.dumps({"classes": META_CLASSES, "colors": PALETTE}))


def scenario(seed=5):
    """Instance 0 = unclassified (category 133), instances 1..7 furniture / wall / appliance / an unlisted category; 1-2 and 4-5 overlap."""
    rng = np.random.Generator(np.random.PCG64(seed))
    cats = {0: 133, 1: 10, 2: 11, 3: 20, 4: 21, 5: 21, 6: 77, 7: 30}
    centres = {1: (0.0, 0.0), 2: (0.8, 0.1), 3: (3.0, 2.0), 4: (6.0, 0.0), 5: (6.7, 0.2), 6: (9.0, 1.0), 7: (12.0, 0.0)}
    pts, ids = [rng.uniform(-1, 1, (150, 3)) * np.array([8.0, 3.0, 0.1]) + np.array([6.0, 0.0, 3.0])], [0] * 150
    for k, (cx, cy) in centres.items():
        n = 70
        yaw = rng.uniform(-0.4, 0.4)
        R = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]])
        pts.append(rng.uniform(-0.5, 0.5, (n, 3)) * np.array([1.2, 0.6, 0.5]) @ R.T + np.array([cx, cy, 0.6]))
        ids += [k] * n
    pts = np.round(np.concatenate(pts), 6) + 0.123456
    ids = np.asarray(ids, dtype=np.int64)
    classes = np.array([cats[i] for i in ids], dtype=np.int64)
    info_pan = [{"id": k, "isthing": k != 0, "category_id": cats[k], "area": int((ids == k).sum()), "name": f"c{cats[k]}"} for k in range(8)]
    info_sem = [{"category_id": int(c), "name": f"c{c}", "area": int((classes == c).sum())} for c in np.unique(classes)]
    return pts, ids, classes, info_pan, info_sem


def main():
    sys.path.insert(0, str(ROOT / "oracle" / "refshim"))
    sys.path.insert(0, str(REF))
    import open3d as o3d   # the shim
    o3d.BOX_MODEL = "pca"
    work = Path(tempfile.mkdtemp(prefix="f3d_master_"))
    try:
        (work / "ref").mkdir()
        shutil.copy(REF / "get3DSeg.py", work / "ref" / "get3DSeg.py")        # throw-away copy OUTSIDE the repo: __file__ decides where the tables are
        write_tables(work)
        spec = importlib.util.spec_from_file_location("ref_get3DSeg", work / "ref" / "get3DSeg.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        pts, ids, classes, info_pan, info_sem = scenario()
        d = work / "scan"
        (d / "panoptic_segmentation").mkdir(parents=True)
        (d / "segmentation").mkdir()
        o3d.io.write_point_cloud(str(d / "panoptic_segmentation" / "pcd.ply"), o3d.geometry.PointCloud(pts))
        np.save(d / "panoptic_segmentation" / "ids.npy", ids)
        np.save(d / "segmentation" / "classes.npy", classes)
        (d / "panoptic_segmentation" / "info.json").write_text(json.dumps(info_pan))
        (d / "segmentation" / "info.json").write_text(json.dumps(info_sem))
        # numpy >= 2 returns a numpy integer from np.count_nonzero (`get3DSeg.py:446`), which json cannot encode; the reference was
        # written against a numpy that returned a Python int.  Teach THIS process's encoder the conversion instead of touching the reference.
        _default = json.JSONEncoder.default
        json.JSONEncoder.default = lambda self, o: int(o) if isinstance(o, np.integer) else _default(self, o)
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                mod.master_classes(d)
        finally:
            json.JSONEncoder.default = _default
        final_pts = o3d.io.read_point_cloud(str(d / "segmentation" / "final_pcd.ply"))
        with open(d / "segmentation" / "final_pcd.ply", "rb") as fp:
            raw = fp.read()
        ncol = len(pts) * 3
        out = {"note": "reference get3DSeg.master_classes run unmodified (box model pca of oracle/refshim/open3d)",
               "class_table": CLASS_TABLE, "palette": PALETTE, "meta_classes": META_CLASSES,
               "points": pts.tolist(), "ids": ids.tolist(), "classes": classes.tolist(), "info_pan": info_pan, "info_sem": info_sem,
               "out_info_pan": json.loads((d / "panoptic_segmentation" / "info.json").read_text()),
               "out_info_sem": json.loads((d / "segmentation" / "info.json").read_text()),
               "out_final_info": json.loads((d / "panoptic_segmentation" / "final_info.json").read_text()),
               "out_ids": np.load(d / "panoptic_segmentation" / "ids.npy").tolist(),
               "out_final_pcd_colors": np.frombuffer(raw[-len(pts) * 27:], dtype=[("x", "<f8"), ("y", "<f8"), ("z", "<f8"), ("r", "u1"), ("g", "u1"), ("b", "u1")])[["r", "g", "b"]].tolist()}
        assert np.allclose(final_pts.points, pts)
        (HERE / "g9_master.json").write_text(json.dumps(out))
        print("final_info ids", [i["id"] for i in out["out_final_info"]], "areas", [i["area"] for i in out["out_final_info"]])
        print("wrote g9_master.json", (HERE / "g9_master.json").stat().st_size, "bytes")
    finally:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
