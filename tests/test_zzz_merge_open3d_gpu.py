"""`check_intersection_open3d` mirror on the GPU against the rows the unmodified reference returned (g7).  Kept in its own
file, last in collection order: the mirror was added after the round's GPU budget was spent, its host logic is checked
without a device in tests/test_merge_golden.py::test_check_intersection_open3d_host_logic."""
import importlib

import numpy as np
import pytest

from test_merge_golden import PKG, _Cloud, _cases


@pytest.mark.gpu
@pytest.mark.parametrize("case,model", list(_cases()))
def test_gpu_check_intersection_open3d(case, model):
    mbb = importlib.import_module(PKG + ".Fusion3DSeg.merge_intersecting_bb")
    pts, ids, info = np.asarray(case["points"]), np.asarray(case["ids"], dtype=np.int64), case["info_sem"]
    id_list = [d["id"] for d in info]
    for id1, ref in case["models"][model]["check_intersection_open3d"].items():
        assert mbb.check_intersection_open3d(int(id1), id_list, ids, pts, _Cloud(pts), info, box_model=model) == ref
