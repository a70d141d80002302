"""`get3DSeg.master_classes` (SURVEY a-16) against the files the UNMODIFIED reference wrote for the same inputs
(tests/golden/make_golden_master.py -> g9_master.json)."""
import importlib
import json
from pathlib import Path

import numpy as np
import pytest

from conftest import PKG_NAME

G9 = json.loads((Path(__file__).parent / "golden" / "g9_master.json").read_text())


def _strip(info):
    return [{k: v for k, v in d.items() if k != "bbox"} for d in info]


def _same_boxes(a, b):
    for da, db in zip(a, b):
        assert (da.get("bbox") is None) == (db.get("bbox") is None)
        if da.get("bbox") is not None:
            ca, cb = np.asarray(da["bbox"]), np.asarray(db["bbox"])
            ca, cb = ca[np.lexsort(np.round(ca, 6).T)], cb[np.lexsort(np.round(cb, 6).T)]
            np.testing.assert_allclose(ca, cb, rtol=0, atol=1e-9)


@pytest.mark.gpu
def test_master_classes_matches_reference_files(engine, tmp_path, monkeypatch):
    g3d = importlib.import_module(PKG_NAME + ".get3DSeg")
    (tmp_path / "classes.csv").write_text("Class_ID,Parent,Parent_ID,flag_infojson,flag_objremoval\n" +
                                          "\n".join(",".join(str(v) for v in row) for row in G9["class_table"]) + "\n")
    (tmp_path / "classes_meta.json").write_text(json.dumps({"classes": G9["meta_classes"], "colors": G9["palette"]}))
    monkeypatch.setattr(g3d, "CLASSES_CSV", tmp_path / "classes.csv")
    monkeypatch.setattr(g3d, "CLASSES_META", tmp_path / "classes_meta.json")
    d = tmp_path / "scan"
    (d / "panoptic_segmentation").mkdir(parents=True)
    (d / "segmentation").mkdir()
    pts = np.asarray(G9["points"])
    g3d.write_ply(d / "panoptic_segmentation" / "pcd.ply", pts)
    np.save(d / "panoptic_segmentation" / "ids.npy", np.asarray(G9["ids"], dtype=np.int64))
    np.save(d / "segmentation" / "classes.npy", np.asarray(G9["classes"], dtype=np.int64))
    (d / "panoptic_segmentation" / "info.json").write_text(json.dumps(G9["info_pan"]))
    (d / "segmentation" / "info.json").write_text(json.dumps(G9["info_sem"]))
    g3d.master_classes(d)
    out_pan = json.loads((d / "panoptic_segmentation" / "info.json").read_text())
    out_sem = json.loads((d / "segmentation" / "info.json").read_text())
    out_fin = json.loads((d / "panoptic_segmentation" / "final_info.json").read_text())
    assert _strip(out_pan) == _strip(G9["out_info_pan"]) and out_sem == G9["out_info_sem"]
    assert _strip(out_fin) == _strip(G9["out_final_info"])          # parent ids, merged areas, surviving instances and their order
    _same_boxes(out_pan, G9["out_info_pan"])
    _same_boxes(out_fin, G9["out_final_info"])
    assert np.array_equal(np.load(d / "panoptic_segmentation" / "ids.npy"), np.asarray(G9["out_ids"]))
    assert np.allclose(g3d.read_ply_points(d / "segmentation" / "final_pcd.ply"), pts)
    with open(d / "segmentation" / "final_pcd.ply", "rb") as fp:
        raw = fp.read()
    rec = np.frombuffer(raw[-len(pts) * 27:], dtype=[("x", "<f8"), ("y", "<f8"), ("z", "<f8"), ("r", "u1"), ("g", "u1"), ("b", "u1")])
    assert np.array_equal(np.stack([rec["r"], rec["g"], rec["b"]], 1), np.asarray(G9["out_final_pcd_colors"], dtype=np.uint8))
