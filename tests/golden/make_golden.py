"""Generate golden input/output vectors by running the UNMODIFIED reference in the build container.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference (`/root/reference`, read-only, absent on the GPU box) is imported as-is; the two third-party
modules it needs that are missing from this image are replaced by `oracle/refshim/{pyquaternion,open3d}`
(semantics documented there).  Outputs are small `.npz` fixtures committed next to this script; the tests
compare both the numpy oracle (`-m "not gpu"`) and the CUDA path (`-m gpu`) against them.
"""
import importlib
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
sys.path.insert(0, str(ROOT))

import cv2  # noqa: E402

scenes = importlib.import_module("3d-point-cloud-segmentation-using-2d-img-segmentation_b200.scenes")
from oracle import f3d_oracle as orc  # noqa: E402  (only for the z-buffer depth synthesis of the inputs)

from Fusion3DSeg import camera_utils as ref_cam  # noqa: E402
from Fusion3DSeg.fusion import Fusion as RefFusion, FrameData as RefFrameData  # noqa: E402
from Fusion3DSeg.intersections import point_inside_polyhedra as ref_pip  # noqa: E402
from Fusion3DSeg.segUtils.voting import VotingSegmentation as RefVoting  # noqa: E402
from RTAB_utils.spatQuad import SpatQuadranion as RefQuat  # noqa: E402
from Fusion3DSeg.segUtils.cv import split_into_instances as ref_split  # noqa: E402


def ref_mod_points(depth_u16, K, wxyz, t):
    """`RTAB2Cache.__getRGBP3d` + `__getModP3d` (`RTAB_utils/ios_rtab.py:164-173,185-191`) -- those are
    private methods of a file-driven class, so their numpy statements are replayed here verbatim in order,
    with the reference's own `SpatQuadranion.rotate`."""
    H, W = depth_u16.shape
    pixel_x, pixel_y = np.meshgrid(np.linspace(0, W - 1, W), np.linspace(0, H - 1, H))
    cx = np.multiply(pixel_x - K[0, 2], depth_u16 / K[0, 0])
    cy = np.multiply(pixel_y - K[1, 2], depth_u16 / K[1, 1])
    org = np.array([cx, cy, depth_u16]).transpose(1, 2, 0).reshape(-1, 3)
    org = np.divide(org, 1000)
    rot = RefQuat(str(wxyz[0]), str(wxyz[1]), str(wxyz[2]), str(wxyz[3]))
    mod = rot.rotate(org) + t
    return org, mod


def ref_level_p_votes(points64, K, w, h, wxyzs, ts, depths, masks, nclasses1, radius, zmin, zmax, max_depth):
    """SURVEY 8(c) level-P composition, every step a reference call."""
    eyes, lookats, spoke_origins, face_normals = RefFusion._get_frustum_data(K, w, h, wxyzs, ts)
    votes = np.zeros((len(points64), nclasses1))
    uv2pts = []
    n_inside = 0
    for j in range(len(ts)):
        plane_pts = np.vstack([spoke_origins[j], (eyes[j] + max_depth * lookats[j])[None, :]])   # fusion.py:254-258
        plane_norms = np.vstack([face_normals[j], (-lookats[j])[None, :]])
        inside = ref_pip(points64, plane_pts, plane_norms)                                       # fusion.py:260
        uv2pt = np.full(h * w, -1, np.int32)
        if inside.any():
            idx = np.where(inside)[0]
            n_inside += len(idx)
            uv = ref_cam.points2pixel(points64[inside], K, wxyzs[j], ts[j])                      # fusion.py:266
            u, v = uv
            ok = (u >= 0) & (u < w) & (v >= 0) & (v < h)
            idx, u, v = idx[ok], u[ok].astype(np.int64), v[ok].astype(np.int64)
            pix = v * w + u
            org, mod = ref_mod_points(depths[j], K, wxyzs[j], ts[j])
            valid = RefFrameData.get_valid(org, zmin, zmax)                                      # fusion.py:62-63
            dist = np.linalg.norm(mod[pix] - points64[idx], axis=-1)                             # fusion.py:224
            vis = valid[pix] & (dist < radius)                                                   # fusion.py:225
            uv2pt[pix[vis]] = idx[vis]
            votes[idx[vis], masks[j].reshape(-1)[pix[vis]]] += 1                                 # voting.py:98
        uv2pts.append(uv2pt)
    return votes, np.stack(uv2pts), n_inside


def main():
    out = {}
    rng = np.random.Generator(np.random.PCG64(20261018))

    # ---- G1: rotate / points2pixel / frustum / cull on random points, three resolutions ----------------
    for tag, (W, H) in {"640": (640, 480), "1920": (1920, 1440), "3840": (3840, 2160)}.items():
        spec = scenes.scaled_spec("C1", npoints=20000, nframes=4, width=W, height=H, seed=rng.integers(1 << 30))
        K = scenes.scaled_intrinsics(W, H)
        wxyz, t = scenes.make_poses(spec)
        pts = scenes.make_cloud(spec)
        p64 = pts.astype(np.float64)
        eyes, lookats, so, fn = RefFusion._get_frustum_data(K, W, H, wxyz, t)
        uvs, insides, rots = [], [], []
        for j in range(len(t)):
            pp = np.vstack([so[j], (eyes[j] + 4.0 * lookats[j])[None, :]])
            pn = np.vstack([fn[j], (-lookats[j])[None, :]])
            insides.append(ref_pip(p64, pp, pn))
            with np.errstate(all="ignore"):
                uvs.append(ref_cam.points2pixel(p64, K, wxyz[j], t[j]))
            rots.append(RefQuat(wxyz[j]).inverse.rotate(p64 - t[j]))
        out[f"g1_{tag}"] = dict(points=pts, K=K, wxyz=wxyz, t=t, W=W, H=H, eyes=eyes, lookats=lookats,
                                face_normals=fn, inside=np.stack(insides), uv=np.stack(uvs),
                                rot=np.stack(rots)[:, :256])

    # ---- G2: level P on a small room, uint16 depth ---------------------------------------------------------
    spec = scenes.scaled_spec("C1", npoints=40000, nframes=8, width=160, height=120, seed=4242)
    W, H = spec.width, spec.height
    K = scenes.scaled_intrinsics(W, H)
    wxyz, t = scenes.make_poses(spec)
    pts = scenes.make_cloud(spec)
    p64 = pts.astype(np.float64)
    eyes, lookats, normals = orc.frustum_data(K, W, H, wxyz, t)
    depths = np.stack([orc.zero_border(orc.zbuffer_splat(p64, K, W, H, wxyz[f], t[f], eyes[f], lookats[f], normals[f],
                                                         4.0), 4) for f in range(spec.nframes)])
    masks = scenes.block_masks((H, W), spec.nframes, seed=spec.seed, block=8)
    votes, uv2pt, n_inside = ref_level_p_votes(p64, K, W, H, wxyz, t, depths, masks, 134, 0.05, 0.1, 4.0, 4.0)
    assert votes.sum() > 1000, votes.sum()
    out["g2_levelp"] = dict(points=pts, K=K, wxyz=wxyz, t=t, W=W, H=H, depths=depths, masks=masks,
                            votes=votes.astype(np.int32), uv2pt=uv2pt, radius=0.05, zmin=0.1, zmax=4.0,
                            max_depth=4.0, n_inside=n_inside)

    # ---- G3: level V through the reference class, from files, with mask resize ---------------------------------
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        (td / "masks").mkdir()
        (td / "uv2pt").mkdir()
        big = scenes.block_masks((2 * H, 2 * W), spec.nframes, seed=77, block=12)
        # make many pixels share a point (the reference's patch merging does) to exercise the de-dup semantics
        coarse = []
        for f in range(spec.nframes):
            u = uv2pt[f].copy()
            grid = (np.arange(H)[:, None] // 5) * 1000 + (np.arange(W)[None, :] // 5)
            pt = (grid * 7 + f * 13) % len(pts)
            u = np.where(rng.random((H, W)) < 0.7, pt, -1).astype(np.int32).reshape(-1)
            coarse.append(u)
            np.save(td / "uv2pt" / f"{f + 1}.npy", u)
            cv2.imwrite(str(td / "masks" / f"{f + 1}.png"), big[f])
        voter = RefVoting(len(pts), (H, W), td / "masks", td / "uv2pt", 133)
        v = voter.vote(resize=True)
        segs = {
            "seg_default": voter.segment(0.5, [86, 114, 115]),
            "seg_all": voter.segment(0.5, None),
            "seg_t075": voter.segment(0.75, None),
            "seg_alias": voter.segment(0.3, [1, 0, 5]),
            "seg_t0": voter.segment(0.0, [3, 2, 1, 0]),
        }
        resized = np.stack([cv2.resize(big[f], (W, H), interpolation=cv2.INTER_NEAREST) for f in range(spec.nframes)])
        out["g3_levelv"] = dict(npts=len(pts), H=H, W=W, uv2pt=np.stack(coarse), masks_big=big, masks_resized=resized,
                                votes=v.astype(np.int32), **segs)
        # segment on the level-P votes as well
        out["g2_levelp"]["seg_default"] = voter.segment(0.5, [86, 114, 115], votes=votes)
        out["g2_levelp"]["seg_all"] = voter.segment(0.5, None, votes=votes)

    # ---- G4: cv2 nearest resize at odd ratios --------------------------------------------------------------------
    rs = {}
    for k, ((sh, sw), (dh, dw)) in enumerate([((37, 53), (20, 31)), ((960, 720), (256, 192)), ((100, 100), (33, 77)),
                                              ((48, 64), (96, 128))]):
        src = rng.integers(0, 134, (sh, sw)).astype(np.uint8)
        rs[f"src{k}"] = src
        rs[f"dst{k}"] = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_NEAREST)
    out["g4_resize"] = rs

    # ---- G5: instance split (connected components per class) through the reference's BFS ---------------------------
    from sklearn.neighbors import KDTree
    import json
    spec5 = scenes.scaled_spec("C1", npoints=6000, nframes=1, seed=99)
    p5 = scenes.make_cloud(spec5).astype(np.float64)
    adj = KDTree(p5).query_radius(p5, r=0.18)                    # as fusion.py:374-375 (r = 2 * radius)
    cell = (np.floor(p5[:, 0] / 1.3) * 7 + np.floor(p5[:, 1] / 1.1) * 3 + np.floor(p5[:, 2] / 1.6)).astype(int)
    classes5 = np.array([86, 114, 115, 133, 5, 86, 133])[cell % 7].astype(np.int64)
    indptr = np.concatenate([[0], np.cumsum([len(a) for a in adj])]).astype(np.int64)
    indices = np.concatenate(adj).astype(np.int64)
    g5 = dict(classes=classes5, indptr=indptr, indices=indices)
    for tag, (inst_cls, minpts) in {"a": ([86, 114, 115], 20), "b": (None, 1), "c": ([115, 86], 100), "d": (None, 50)}.items():
        insts, ids5, info5, cls5 = ref_split(classes5, adj, 133, inst_cls, minpts)
        g5[f"ids_{tag}"] = ids5
        g5[f"classes_{tag}"] = cls5
        g5[f"ninst_{tag}"] = len(insts)
        g5[f"info_{tag}"] = json.dumps(info5)
    out["g5_instances"] = g5

    for name, d in out.items():
        np.savez_compressed(HERE / f"{name}.npz", **d)
        print(name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in d.items()})


if __name__ == "__main__":
    main()
