"""Instance split: oracle vs the reference's BFS output (CPU), GPU mirror vs the same vectors, and the
full `get3DSeg.segment` flow through split -> panoptic dump -> master_classes -> merge_bb (GPU)."""
import importlib
import json
import pickle

import numpy as np
import pytest

from conftest import PKG_NAME, load_golden
from oracle import f3d_oracle as orc

CASES = {"a": ([86, 114, 115], 20), "b": (None, 1), "c": ([115, 86], 100), "d": (None, 50)}


@pytest.mark.parametrize("tag", sorted(CASES))
def test_oracle_split_matches_reference(tag):
    g = load_golden("g5_instances")
    ic, mp = CASES[tag]
    n, ids, info, cls = orc.split_into_instances(g["classes"], g["indptr"], g["indices"], 133, ic, mp)
    assert n == int(g[f"ninst_{tag}"]) and np.array_equal(ids, g[f"ids_{tag}"]) and np.array_equal(cls, g[f"classes_{tag}"])
    assert info == json.loads(str(g[f"info_{tag}"]))


def _numpy_union_find(n, edges):
    """Stand-in with `engine.union_find`'s contract: label = smallest index of the connected component."""
    import torch
    parent = np.arange(n)

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x
    for a, b in edges.numpy().astype(np.int64):
        ra, rb = find(a), find(b)
        if ra != rb:
            parent[max(ra, rb)] = min(ra, rb)
    return torch.as_tensor(np.array([find(i) for i in range(n)], dtype=np.int32))


@pytest.mark.parametrize("tag", sorted(CASES))
def test_split_host_logic_matches_reference(tag, monkeypatch):
    """The mirror's HOST side (edge filtering, component table, the reference's id numbering and small-component fold) with the
    GPU union-find replaced by a numpy stand-in of the same contract -- runs without a device; the kernel itself is
    checked on the GPU below."""
    import torch
    cv = importlib.import_module(PKG_NAME + ".Fusion3DSeg.segUtils.cv")
    monkeypatch.setattr(cv, "require_cuda", lambda: torch.device("cpu"))
    monkeypatch.setattr(cv.engine, "union_find", _numpy_union_find)
    g = load_golden("g5_instances")
    ic, mp = CASES[tag]
    adj = [g["indices"][g["indptr"][i]:g["indptr"][i + 1]] for i in range(len(g["classes"]))]
    for a in (adj, (g["indptr"], g["indices"])):                    # the reference's list form and the CSR pair
        insts, ids, info, cls = cv.split_into_instances(g["classes"], a, 133, ic, mp)
        assert len(insts) == int(g[f"ninst_{tag}"]) and np.array_equal(ids, g[f"ids_{tag}"])
        assert np.array_equal(cls, g[f"classes_{tag}"]) and info == json.loads(str(g[f"info_{tag}"]))


@pytest.mark.gpu
@pytest.mark.parametrize("tag", sorted(CASES))
def test_gpu_split_matches_reference(engine, tag):
    cv = importlib.import_module(PKG_NAME + ".Fusion3DSeg.segUtils.cv")
    g = load_golden("g5_instances")
    ic, mp = CASES[tag]
    adj = [g["indices"][g["indptr"][i]:g["indptr"][i + 1]] for i in range(len(g["classes"]))]   # the reference's list form
    insts, ids, info, cls = cv.split_into_instances(g["classes"], adj, 133, ic, mp)
    assert len(insts) == int(g[f"ninst_{tag}"]) and np.array_equal(ids, g[f"ids_{tag}"])
    assert np.array_equal(cls, g[f"classes_{tag}"]) and info == json.loads(str(g[f"info_{tag}"]))
    # CSR form and a denser random graph against the oracle
    rng = np.random.default_rng(3)
    n = 3000
    src = rng.integers(0, n, 9000)
    dst = np.clip(src + rng.integers(-40, 41, 9000), 0, n - 1)
    rows = [[] for _ in range(n)]
    for a, b in zip(src, dst):
        rows[a].append(b)
        rows[b].append(a)
    indptr = np.concatenate([[0], np.cumsum([len(r) for r in rows])])
    indices = np.array([x for r in rows for x in r], dtype=np.int64)
    classes = rng.choice([86, 114, 115, 133, 7], n)
    for ic2, mp2 in [([86, 114, 115], 5), (None, 3), ([7], 1)]:
        on, oids, oinfo, ocls = orc.split_into_instances(classes, indptr, indices, 133, ic2, mp2)
        insts, ids, info, cls = cv.split_into_instances(classes, (indptr, indices), 133, ic2, mp2)
        assert len(insts) == on and np.array_equal(ids, oids) and np.array_equal(cls, ocls) and info == oinfo


@pytest.mark.gpu
def test_get3dseg_full_flow_with_adjacency(engine, tmp_path, monkeypatch):
    """segment() with an adjacency list: votes -> classes -> instances -> panoptic dumps -> parents -> merged boxes."""
    import cv2
    from sklearn.neighbors import KDTree
    g3 = importlib.import_module(PKG_NAME + ".get3DSeg")
    scenes = importlib.import_module(PKG_NAME + ".scenes")
    spec = scenes.scaled_spec("C1", npoints=8000, nframes=1, seed=5)
    pts = scenes.make_cloud(spec).astype(np.float64)
    N, H, W, F = len(pts), 60, 80, 6
    rng = np.random.default_rng(8)
    # level-V inputs built so that spatial blobs get consistent classes (pixel p of frame f sees point (p*7+f) % N)
    blob = (np.floor(pts[:, 0] / 2.0) * 5 + np.floor(pts[:, 1] / 2.0)).astype(int)
    point_class = np.array([86, 114, 115, 20])[blob % 4]
    (tmp_path / "fusion" / "uv2pt").mkdir(parents=True)
    (tmp_path / "masks").mkdir()
    for f in range(F):
        uv = ((np.arange(H * W) * 7 + f * 1013) % N).astype(np.int32)
        uv[rng.random(H * W) < 0.2] = -1
        m = np.where(uv >= 0, point_class[np.maximum(uv, 0)], 133).astype(np.uint8).reshape(H, W)
        np.save(tmp_path / "fusion" / "uv2pt" / f"{f}.npy", uv)
        cv2.imwrite(str(tmp_path / "masks" / f"{f}.png"), m)
    adj = KDTree(pts).query_radius(pts, r=0.25)
    with open(tmp_path / "fusion" / "fusion_data.pkl", "wb") as fp:
        pickle.dump({"points": pts, "normals": np.zeros_like(pts), "colors": np.zeros_like(pts), "nmerges": None,
                     "occurences": None, "nframes": F, "depth_hw": (H, W)}, fp)
    with open(tmp_path / "fusion" / "adj.pkl", "wb") as fp:
        pickle.dump(np.array(adj, dtype=object), fp)
    csv = tmp_path / "classes.csv"
    csv.write_text("Class_ID,Parent,Parent_ID,flag_infojson,flag_objremoval\n86,door,1,1,0\n114,window,2,1,0\n115,wall,3,1,0\n"
                   "133,unclassified,0,1,1\n")
    meta = tmp_path / "classes_meta.json"
    meta.write_text(json.dumps({"classes": ["unclassified", "door", "window", "wall"],
                                "colors": [[0, 0, 0], [255, 0, 0], [0, 255, 0], [0, 0, 255]]}))
    monkeypatch.setattr(g3, "CLASSES_CSV", csv)
    monkeypatch.setattr(g3, "CLASSES_META", meta)
    assert g3.segment(tmp_path, tmp_path / "masks", min_pts_per_inst=30, verbose=False) is None   # reference returns None here
    classes = np.load(tmp_path / "segmentation" / "classes.npy")
    votes = np.load(tmp_path / "segmentation" / "votes.npy")
    # semantic part equals the oracle
    ov = np.zeros((N, 134), np.int64)
    for f in range(F):
        orc.vote_uv2pt(ov, np.load(tmp_path / "fusion" / "uv2pt" / f"{f}.npy"), cv2.imread(str(tmp_path / "masks" / f"{f}.png"), 0))
    assert np.array_equal(votes, ov.astype(np.float64))
    assert np.array_equal(classes, orc.segment(ov, 133, 0.5, [86, 114, 115]))
    # panoptic part: ids written before the merge equal the oracle's split; merge only ever relabels to an existing id
    indptr = np.concatenate([[0], np.cumsum([len(a) for a in adj])])
    on, oids, oinfo, ocls = orc.split_into_instances(classes, indptr, np.concatenate(adj), 133, [86, 114, 115], 30)
    final_ids = np.load(tmp_path / "panoptic_segmentation" / "ids.npy")
    assert on > 3 and set(np.unique(final_ids)) <= set(np.unique(oids))
    changed = final_ids != oids
    assert np.all(np.isin(oids[changed], np.unique(oids)))                        # merged instances vanish as whole units
    for k in np.unique(oids[changed]):
        assert len(np.unique(final_ids[oids == k])) == 1
    info = json.load(open(tmp_path / "panoptic_segmentation" / "final_info.json"))
    assert all("parent_id" in d for d in info) and (tmp_path / "segmentation" / "final_pcd.ply").is_file()
    pan = json.load(open(tmp_path / "panoptic_segmentation" / "info.json"))
    assert [d["area"] for d in pan if d["category_id"] != 133 or True][:0] == []   # file is valid json with area fields
    assert sum(d["area"] for d in oinfo) == N
