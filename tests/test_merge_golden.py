"""Box-merge parity against vectors the UNMODIFIED reference `merge_intersecting_bb.py` produced
(tests/golden/make_golden_merge.py -> g7_merge.json): oracle restatement on CPU, product `merge_bb` / `cal_min_max` /
`check_intersection` on the GPU."""
import copy
import importlib
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import f3d_oracle as orc

G7 = json.loads((Path(__file__).parent / "golden" / "g7_merge.json").read_text())
PKG = "3d-point-cloud-segmentation-using-2d-img-segmentation_b200"


def _cases():
    for case in G7["cases"]:
        for model in ("pca", "aabb"):
            yield pytest.param(case, model, id=f"{case['kind']}{case['seed']}-{model}")


def _strip(info):
    return [{k: v for k, v in d.items() if k != "bbox"} for d in info]


def _same_boxes(a, b):
    """Corner sets equal (the corner ORDER depends on eigenvector signs, which LAPACK / Jacobi may choose differently)."""
    for da, db in zip(a, b):
        assert ("bbox" in da) == ("bbox" in db)
        if "bbox" in da:
            ca, cb = np.asarray(da["bbox"]), np.asarray(db["bbox"])
            ca, cb = ca[np.lexsort(np.round(ca, 6).T)], cb[np.lexsort(np.round(cb, 6).T)]
            np.testing.assert_allclose(ca, cb, rtol=0, atol=1e-9)


@pytest.mark.parametrize("case,model", list(_cases()))
def test_oracle_merge_bb_matches_reference(case, model):
    pts, ids, info = np.asarray(case["points"]), np.asarray(case["ids"], dtype=np.int64), copy.deepcopy(case["info_sem"])
    gold = case["models"][model]
    info, ids = orc.merge_bb_sequential(info, ids, orc.merge_hit_fn(pts, model))
    info = orc.merge_bb_final_boxes(info, ids, pts, model)
    assert np.array_equal(ids, np.asarray(gold["final_ids"]))
    assert _strip(info) == _strip(gold["final_info"])          # surviving ids, order, areas (incl. the position-vs-id quirk)
    _same_boxes(info, gold["final_info"])


@pytest.mark.parametrize("case,model", list(_cases()))
def test_oracle_cal_min_max_and_check_intersection(case, model):
    pts, ids, info = np.asarray(case["points"]), np.asarray(case["ids"], dtype=np.int64), case["info_sem"]
    gold = case["models"][model]
    for k, ref in gold["cal_min_max"].items():
        got = orc.cal_min_max(orc.box_corners(*orc.fit_box(pts[ids == int(k)], model)))
        for g, r in zip(got, ref):
            np.testing.assert_allclose(g, np.asarray(r), rtol=0, atol=1e-12)
    id_list = [d["id"] for d in info]
    assert gold["check_intersection"], "fixture has no check_intersection rows"
    for id1, ref in gold["check_intersection"].items():
        assert orc.check_intersection_as_shipped(int(id1), id_list, ids, pts, info, model) == ref


class _Cloud:
    def __init__(self, pts):
        self.points = pts


@pytest.mark.gpu
@pytest.mark.parametrize("case,model", list(_cases()))
def test_gpu_merge_bb_matches_reference(case, model, tmp_path):
    mbb = importlib.import_module(PKG + ".Fusion3DSeg.merge_intersecting_bb")
    pts, ids, info = np.asarray(case["points"]), np.asarray(case["ids"], dtype=np.int64), copy.deepcopy(case["info_sem"])
    gold = case["models"][model]
    (tmp_path / "panoptic_segmentation").mkdir()
    mbb.merge_bb(tmp_path, info, ids, _Cloud(pts), box_model=model)
    assert np.array_equal(ids, np.asarray(gold["final_ids"]))
    assert _strip(info) == _strip(gold["final_info"])
    _same_boxes(info, gold["final_info"])
    assert np.array_equal(np.load(tmp_path / "panoptic_segmentation" / "ids.npy"), ids)
    assert _strip(json.loads((tmp_path / "panoptic_segmentation" / "final_info.json").read_text())) == _strip(gold["final_info"])


@pytest.mark.gpu
@pytest.mark.parametrize("case,model", list(_cases()))
def test_gpu_cal_min_max_and_check_intersection(case, model):
    mbb = importlib.import_module(PKG + ".Fusion3DSeg.merge_intersecting_bb")
    pts, ids, info = np.asarray(case["points"]), np.asarray(case["ids"], dtype=np.int64), case["info_sem"]
    gold = case["models"][model]
    for k, ref in gold["cal_min_max"].items():
        got = mbb.cal_min_max(int(k), ids, pts, box_model=model)
        for g, r in zip(got, ref):
            np.testing.assert_allclose(g, np.asarray(r), rtol=0, atol=1e-9)
    id_list = [d["id"] for d in info]
    for id1, ref in gold["check_intersection"].items():
        assert mbb.check_intersection(int(id1), id_list, ids, pts, info, box_model=model) == ref


@pytest.mark.parametrize("case,model", list(_cases()))
def test_oracle_check_intersection_open3d(case, model):
    """The oracle's hit function is the geometric part of `check_intersection_open3d`; wrapped in the loop as shipped it must
    give the rows the unmodified reference returned."""
    pts, ids, info = np.asarray(case["points"]), np.asarray(case["ids"], dtype=np.int64), case["info_sem"]
    hit = orc.merge_hit_fn(pts, model)
    for id1s, ref in case["models"][model]["check_intersection_open3d"].items():
        id1, got = int(id1s), []
        if hit(id1, None, ids) is not False:
            for id2 in range(1, len(info)):
                if id1 != id2 and id2 < len(info) - 1 and id1 < len(info) - 1 and info[id1]["parent_id"] == info[id2]["parent_id"]:
                    h = hit(id1, id2, ids)
                    if h is None:
                        break
                    if h:
                        got.append(id2)
        assert got == ref


def _device_free(monkeypatch):
    """The mirror module with the two GPU operators it composes (`engine.obb_fit`, `engine.obb_contains`) replaced by numpy
    stand-ins of the same contract: what remains under test is the mirror's HOST logic.  The operators themselves are checked
    on the GPU (test_gpu_* here, tests/test_gpu_round2.py)."""
    import torch
    mbb = importlib.import_module(PKG + ".Fusion3DSeg.merge_intersecting_bb")

    def fake_obb_fit(points64, ids, instance_ids, model_="pca"):
        p, i = points64.numpy(), ids.numpy()
        boxes, counts = np.zeros((len(instance_ids), 15)), np.zeros(len(instance_ids), dtype=np.int64)
        for k, inst in enumerate(instance_ids):
            sel = i == inst
            counts[k] = sel.sum()
            if counts[k] >= 1:
                c, R, e = orc.fit_box(p[sel], model_)
                boxes[k] = np.concatenate([c, R.reshape(-1), e])
        return torch.as_tensor(boxes), torch.as_tensor(counts)

    def fake_obb_contains(points64, boxes15):
        p, b = points64.numpy(), boxes15.numpy().reshape(-1, 15)
        return torch.as_tensor(np.stack([orc.obb_contains(r[:3], r[3:12].reshape(3, 3), r[12:], p) for r in b]).astype(np.uint8))

    monkeypatch.setattr(mbb, "require_cuda", lambda: torch.device("cpu"))
    monkeypatch.setattr(mbb.engine, "obb_fit", fake_obb_fit)
    monkeypatch.setattr(mbb.engine, "obb_contains", fake_obb_contains)
    return mbb


@pytest.mark.parametrize("case,model", list(_cases()))
def test_check_intersection_open3d_host_logic(case, model, monkeypatch):
    """Guards, early return and candidate batching of the mirror, without a device."""
    mbb = _device_free(monkeypatch)
    pts, ids, info = np.asarray(case["points"]), np.asarray(case["ids"], dtype=np.int64), case["info_sem"]
    id_list = [d["id"] for d in info]
    for id1, ref in case["models"][model]["check_intersection_open3d"].items():
        assert mbb.check_intersection_open3d(int(id1), id_list, ids, pts, _Cloud(pts), info, box_model=model) == ref


@pytest.mark.parametrize("case,model", list(_cases()))
def test_merge_bb_host_logic(case, model, monkeypatch, tmp_path):
    """The sequential driver of the mirror `merge_bb` (membership cache, re-fit after a relabel, index-as-id and shrinking-list
    quirks, the two output files) against the unmodified reference's result, without a device."""
    mbb = _device_free(monkeypatch)
    pts, ids, info = np.asarray(case["points"]), np.asarray(case["ids"], dtype=np.int64), copy.deepcopy(case["info_sem"])
    gold = case["models"][model]
    mbb.merge_bb(tmp_path, info, ids, _Cloud(pts), box_model=model)
    assert np.array_equal(ids, np.asarray(gold["final_ids"]))
    assert _strip(info) == _strip(gold["final_info"])
    _same_boxes(info, gold["final_info"])
    assert np.array_equal(np.load(tmp_path / "panoptic_segmentation" / "ids.npy"), ids)


@pytest.mark.parametrize("case,model", list(_cases()))
def test_cal_min_max_and_check_intersection_host_logic(case, model, monkeypatch):
    mbb = _device_free(monkeypatch)
    pts, ids, info = np.asarray(case["points"]), np.asarray(case["ids"], dtype=np.int64), case["info_sem"]
    gold = case["models"][model]
    for k, ref in gold["cal_min_max"].items():
        for g, r in zip(mbb.cal_min_max(int(k), ids, pts, box_model=model), ref):
            np.testing.assert_allclose(g, np.asarray(r), rtol=0, atol=1e-9)
    id_list = [d["id"] for d in info]
    for id1, ref in gold["check_intersection"].items():
        assert mbb.check_intersection(int(id1), id_list, ids, pts, info, box_model=model) == ref
