"""CPU ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy restatement of the reference's multi-view 2D->3D label-fusion path.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import it, and only as
the checker / the CPU arm; the product (`3d-point-cloud-segmentation-using-2d-img-segmentation_b200/`) never
does and fails loudly when its CUDA library is missing.

Every function cites the reference file:line it restates (paths relative to the reference repository root).
Floating-point evaluation order is *fixed* here (plain IEEE-754 binary64 multiplies / adds / divides / square
roots, left to right as written, never fused), because the reference itself leaves it to its BLAS
(`K @ P.T`, `np.dot`, `einsum`).  The CUDA exact path evaluates the identical sequence with `__dmul_rn` /
`__dadd_rn` / `__ddiv_rn` / `__dsqrt_rn`, so integer outcomes (pixels, votes, labels) are bit-exact against
this file.  Pinning: `tests/golden/make_golden.py` runs the UNMODIFIED reference functions (imported from
`/root/reference` in the build container with the `oracle/refshim` stand-ins for the absent pyquaternion /
open3d) on seeded scenes and stores their outputs under `tests/golden/`; `tests/test_oracle_golden.py` checks
this restatement against those vectors (`tests/golden/make_golden_ingest.py` does the same for the scan-directory
readers, which are host code of the product and are compared with the reference's readers directly).  The box-merge part has no runnable reference in this image
(Open3D absent) -> "parity unpinned" for OBB fitting; the pair predicate and merge drivers are restated from
the source text only.  UPDATE (round 2): the merge DRIVERS are pinned too -- `tests/golden/make_golden_merge.py` runs the
unmodified `Fusion3DSeg/merge_intersecting_bb.py` (`merge_bb`, `check_intersection_open3d`, `update_id_info`, `cal_min_max`,
`check_intersection`) on the stated box models of `oracle/refshim/open3d` and `tests/test_merge_golden.py` checks `fit_box`,
`cal_min_max`, `check_intersection_as_shipped`, `merge_bb_sequential` against those outputs.  Still unpinned: Open3D's
hull-based box FIT itself (`create_from_points`), which needs Open3D.
"""
from __future__ import annotations

import numpy as np

# ----------------------------------------------------------------------------------------------------------
# quaternion arithmetic
# ----------------------------------------------------------------------------------------------------------


def quat_inverse(q):
    """pyquaternion `Quaternion.inverse` = conjugate / sum-of-squares, NO normalisation
    (third-party, call site `Fusion3DSeg/camera_utils.py:22`).  q = (w, x, y, z) float64.
    Order fixed here: ss = ((w*w + x*x) + y*y) + z*z."""
    q = np.asarray(q, dtype=np.float64)
    ss = ((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3]
    return np.array([q[0] / ss, -q[1] / ss, -q[2] / ss, -q[3] / ss], dtype=np.float64)


def _cross(a0, a1, a2, b0, b1, b2):
    """np.cross component formulas: two rounded products and one subtraction each."""
    return a1 * b2 - a2 * b1, a2 * b0 - a0 * b2, a0 * b1 - a1 * b0


def quat_rotate(q, p):
    """`SpatQuadranion.rotate` (`RTAB_utils/spatQuad.py:6-28`): raw Hamilton sandwich q p q* on the
    UN-normalised elements of q.  p: [N,3] float64 -> [N,3] float64."""
    q = np.asarray(q, dtype=np.float64)
    p = np.asarray(p, dtype=np.float64)
    rq, v0, v1, v2 = q[0], q[1], q[2], q[3]
    n0, n1, n2 = -v0, -v1, -v2                                   # vq_ = -vq            spatQuad.py:18
    p0, p1, p2 = p[:, 0], p[:, 1], p[:, 2]
    rqp = -((p0 * v0 + p1 * v1) + p2 * v2)                       # -np.dot(p, vq)       spatQuad.py:22
    c0, c1, c2 = _cross(v0, v1, v2, p0, p1, p2)                  # np.cross(vq, p)      spatQuad.py:23
    w0, w1, w2 = rq * p0 + c0, rq * p1 + c1, rq * p2 + c2        # vqp                  spatQuad.py:23
    d0, d1, d2 = _cross(w0, w1, w2, n0, n1, n2)                  # np.cross(vqp, vq_)   spatQuad.py:27
    o0 = (rqp * n0 + rq * w0) + d0                               # vqpq                 spatQuad.py:27
    o1 = (rqp * n1 + rq * w1) + d1
    o2 = (rqp * n2 + rq * w2) + d2
    return np.stack([o0, o1, o2], axis=1)


# ----------------------------------------------------------------------------------------------------------
# projection  (a-1)
# ----------------------------------------------------------------------------------------------------------


def project_homogeneous(points, K, quat, t):
    """First three statements of `points2pixel` (`Fusion3DSeg/camera_utils.py:21-23`): returns the three rows
    of `K @ P.T` as (h0, h1, h2).  Row i = (K[i,0]*X + K[i,1]*Y) + K[i,2]*Z."""
    points = np.asarray(points, dtype=np.float64)
    K = np.asarray(K, dtype=np.float64)
    t = np.asarray(t, dtype=np.float64)
    P = points - t[None, :]                                      # camera_utils.py:21
    P = quat_rotate(quat_inverse(quat), P)                       # camera_utils.py:22
    X, Y, Z = P[:, 0], P[:, 1], P[:, 2]
    h = [(K[i, 0] * X + K[i, 1] * Y) + K[i, 2] * Z for i in range(3)]   # camera_utils.py:23
    return h[0], h[1], h[2]


def points2pixel(points, K, quat, t):
    """`points2pixel` (`Fusion3DSeg/camera_utils.py:9-26`): world -> int32 [2,N] pixel (u row 0, v row 1)."""
    h0, h1, h2 = project_homogeneous(points, K, quat, t)
    with np.errstate(divide="ignore", invalid="ignore"):
        u = h0 / h2                                              # camera_utils.py:24
        v = h1 / h2
        uv = np.floor(np.stack([u, v], axis=0)).astype(np.int32)  # camera_utils.py:25
    return uv


# ----------------------------------------------------------------------------------------------------------
# frustum set-up  (a-3) and cull (a-4)
# ----------------------------------------------------------------------------------------------------------


def intrinsic_inverse(K):
    """`np.linalg.inv(K)` (`camera_utils.py:73`) restated in closed form for an upper-triangular pin-hole
    intrinsic matrix (the only form the reference produces, `RTAB_utils/ios_rtab.py:125-131`)."""
    K = np.asarray(K, dtype=np.float64)
    if K[1, 0] != 0 or K[2, 0] != 0 or K[2, 1] != 0:
        raise ValueError("intrinsic matrix must be upper triangular")
    i00, i11, i22 = 1.0 / K[0, 0], 1.0 / K[1, 1], 1.0 / K[2, 2]
    i01 = -((K[0, 1] * i11) * i00)
    i12 = -((K[1, 2] * i22) * i11)
    i02 = -((K[0, 1] * i12 + K[0, 2] * i22) * i00)
    return np.array([[i00, i01, i02], [0.0, i11, i12], [0.0, 0.0, i22]], dtype=np.float64)


def frustum_data(K, w, h, wxyzs, ts):
    """`Fusion._get_frustum_data` (`Fusion3DSeg/fusion.py:119-132`) = `get_camera_frustum`
    (`camera_utils.py:60-93`) + `camera2world` (`:96-132`) + `get_frustum_unit_vectors` (`:135-150`) +
    `get_frustum_face_normals` (`:153-171`).  Returns eyes [F,3], lookats [F,3], face_normals [F,4,3]."""
    Ki = intrinsic_inverse(K)
    wxyzs = np.asarray(wxyzs, dtype=np.float64).reshape(-1, 4)
    ts = np.asarray(ts, dtype=np.float64).reshape(-1, 3)
    pix = np.array([[0, 0, 0], [0, 0, 1], [w, 0, 1], [w, h, 1], [0, h, 1], [w / 2, h / 2, 1]], dtype=np.float64)
    cam = np.stack([(Ki[i, 0] * pix[:, 0] + Ki[i, 1] * pix[:, 1]) + Ki[i, 2] * pix[:, 2] for i in range(3)], axis=1)
    F = len(ts)
    eyes = np.zeros((F, 3))
    lookats = np.zeros((F, 3))
    normals = np.zeros((F, 4, 3))
    for f in range(F):
        world = quat_rotate(wxyzs[f], cam) + ts[f][None, :]      # camera_utils.py:128-129
        eye = world[0]
        vec = world[1:] - eye[None, :]                           # camera_utils.py:147
        nrm = np.sqrt((vec[:, 0] * vec[:, 0] + vec[:, 1] * vec[:, 1]) + vec[:, 2] * vec[:, 2])
        dirs = vec / nrm[:, None]                                # camera_utils.py:148
        a = world[1:5]                                           # four corners, fusion.py:124
        b = np.roll(a, -1, axis=0)                               # camera_utils.py:164-166
        ea, eb = a - eye[None, :], b - eye[None, :]
        n0, n1, n2 = _cross(ea[:, 0], ea[:, 1], ea[:, 2], eb[:, 0], eb[:, 1], eb[:, 2])
        nn = np.sqrt((n0 * n0 + n1 * n1) + n2 * n2)
        normals[f] = np.stack([n0 / nn, n1 / nn, n2 / nn], axis=1)
        eyes[f] = eye
        lookats[f] = dirs[4]
    return eyes, lookats, normals


def frame_planes(eye, lookat, face_normals, max_depth):
    """Plane set handed to the cull for one frame (`Fusion3DSeg/fusion.py:254-258`): four side faces through
    the eye + a far plane at eye + max_depth * lookat with normal -lookat.  -> ([5,3], [5,3])."""
    far_pt = eye + max_depth * lookat
    pts = np.vstack([np.repeat(eye[None, :], 4, axis=0), far_pt[None, :]])
    nrm = np.vstack([face_normals, -lookat[None, :]])
    return pts, nrm


def point_inside_polyhedra(points, plane_points, normals):
    """`point_inside_polyhedra` (`Fusion3DSeg/intersections.py:146-164`): inside <=> every
    dp_m = ((p-a_m)_0*n_m0 + (p-a_m)_1*n_m1) + (p-a_m)_2*n_m2 >= 0."""
    points = np.asarray(points, dtype=np.float64)
    inside = np.ones(len(points), dtype=bool)
    for a, n in zip(np.asarray(plane_points, dtype=np.float64), np.asarray(normals, dtype=np.float64)):
        d0, d1, d2 = points[:, 0] - a[0], points[:, 1] - a[1], points[:, 2] - a[2]
        dp = (d0 * n[0] + d1 * n[1]) + d2 * n[2]
        inside &= dp >= 0
    return inside


# ----------------------------------------------------------------------------------------------------------
# per-frame depth data contract  (a-5)
# ----------------------------------------------------------------------------------------------------------

DEPTH_U16_MM = 0   # uint16 millimetres, the RTAB export convention (ios_rtab.py:97-113,185)
DEPTH_F32_M = 1    # float32 metres -- extension (no /1000 step); documented in DESIGN.md


def depth_to_points(depth, K, quat, t, depth_fmt=DEPTH_U16_MM, pix=None):
    """Back-projection of a depth image to camera-space `orgPoints` (metres) and world-space `modPoints`
    (`RTAB_utils/ios_rtab.py:164-173,185-191`): X = (px-cx)*(d/fx), Y = (py-cy)*(d/fy), Z = d, then /1000,
    then forward quaternion rotate (un-normalised) + translation.  `pix` optionally restricts to a flat pixel
    index subset (v*W+u).  depth: [H,W]."""
    K = np.asarray(K, dtype=np.float64)
    H, W = depth.shape
    d = np.asarray(depth).reshape(-1).astype(np.float64)
    if pix is None:
        pix = np.arange(H * W)
    pix = np.asarray(pix, dtype=np.int64)
    d = d[pix]
    px = (pix % W).astype(np.float64)
    py = (pix // W).astype(np.float64)
    X = (px - K[0, 2]) * (d / K[0, 0])                           # ios_rtab.py:168
    Y = (py - K[1, 2]) * (d / K[1, 1])                           # ios_rtab.py:169
    Z = d
    if depth_fmt == DEPTH_U16_MM:
        X, Y, Z = X / 1000, Y / 1000, Z / 1000                   # ios_rtab.py:185
    org = np.stack([X, Y, Z], axis=1)
    mod = quat_rotate(quat, org) + np.asarray(t, dtype=np.float64)[None, :]   # ios_rtab.py:189-190
    return org, mod


def get_valid(org_points, mindist, maxdist):
    """`FrameData.get_valid` (`Fusion3DSeg/fusion.py:50-64`)."""
    z = org_points[:, 2]
    return (z > mindist) & (z <= maxdist)


# ----------------------------------------------------------------------------------------------------------
# level P: fixed-cloud project + visibility + gather + vote  (SURVEY 8c composition)
# ----------------------------------------------------------------------------------------------------------


def fuse_frame_visibility(points, K, w, h, quat, t, eye, lookat, face_normals, depth, depth_fmt,
                          radius, zmin, zmax, max_depth, stats=None):
    """One frame of the level-P composition.  Returns (idx, pix): indices of the cloud points that frame
    *sees* and the flat pixel each one lands in.

    (1) cull: `point_inside_polyhedra` with the 5 planes (`fusion.py:254-260`);
    (2) `points2pixel` on survivors (`fusion.py:266`);
    (3) drop u not in [0,w) / v not in [0,h) (`fuse` is immune through slice clamping, `fusion.py:274-279`);
    (4) visibility = depth-valid pixel (`fusion.py:62-63`) AND single-pixel `criterion`
        ||modPoints[pix] - p|| < radius (`fusion.py:223-225`, stride 1 => half 0, `fusion.py:232,274-277`).
    """
    points = np.asarray(points, dtype=np.float64)
    ppts, pnrm = frame_planes(eye, lookat, face_normals, max_depth)
    inside = point_inside_polyhedra(points, ppts, pnrm)
    idx = np.nonzero(inside)[0]
    if len(idx) == 0:
        return idx, idx
    h0, h1, h2 = project_homogeneous(points[idx], K, quat, t)
    with np.errstate(divide="ignore", invalid="ignore"):
        uf, vf = h0 / h2, h1 / h2
    u = np.floor(uf)
    v = np.floor(vf)
    ok = (u >= 0) & (u < w) & (v >= 0) & (v < h)
    if stats is not None:
        fu, fv = uf[ok] - u[ok], vf[ok] - v[ok]
        stats["in_bounds"] = stats.get("in_bounds", 0) + int(ok.sum())
        stats["near_edge_1e-4"] = stats.get("near_edge_1e-4", 0) + int(
            ((np.minimum(fu, 1 - fu) < 1e-4) | (np.minimum(fv, 1 - fv) < 1e-4)).sum())
    idx = idx[ok]
    pix = (v[ok].astype(np.int64) * w + u[ok].astype(np.int64))
    org, mod = depth_to_points(depth, K, quat, t, depth_fmt, pix)
    valid = get_valid(org, zmin, zmax)
    diff = mod - points[idx]
    dist = np.sqrt((diff[:, 0] * diff[:, 0] + diff[:, 1] * diff[:, 1]) + diff[:, 2] * diff[:, 2])
    vis = valid & (dist < radius)                                # fusion.py:225 (strict)
    return idx[vis], pix[vis]


def fuse_project_vote(points, K, w, h, wxyzs, ts, depths, masks, nclasses1, depth_fmt=DEPTH_U16_MM,
                      radius=0.05, zmin=0.1, zmax=4.0, max_depth=4.0, votes=None, stats=None, frames=None):
    """Level-P oracle: for every frame accumulate `votes[idx, mask[pix]] += 1` (`segUtils/voting.py:98`).
    votes: int64 [N, nclasses1]."""
    points = np.asarray(points, dtype=np.float64)
    N = len(points)
    if votes is None:
        votes = np.zeros((N, nclasses1), dtype=np.int64)
    eyes, lookats, normals = frustum_data(K, w, h, wxyzs, ts)
    frames = range(len(ts)) if frames is None else frames
    for f in frames:
        idx, pix = fuse_frame_visibility(points, K, w, h, wxyzs[f], ts[f], eyes[f], lookats[f], normals[f],
                                         depths[f], depth_fmt, radius, zmin, zmax, max_depth, stats)
        if len(idx):
            cls = np.asarray(masks[f]).reshape(-1)[pix].astype(np.int64)
            votes[idx, cls] += 1        # one pixel per point per frame => no duplicates to collapse
    return votes


def frame_uv2pt(points, K, w, h, quat, t, eye, lookat, face_normals, depth, depth_fmt, radius, zmin, zmax,
                max_depth):
    """The level-P association written in the reference's exchange format (`fusion.py:253,297,322`):
    int32 [h*w], value = LAST (highest-index) cloud point seen through that pixel, -1 = none.  numpy's
    `uv2pt[pix] = idx` keeps the last write; idx is ascending."""
    idx, pix = fuse_frame_visibility(points, K, w, h, quat, t, eye, lookat, face_normals, depth, depth_fmt,
                                     radius, zmin, zmax, max_depth)
    out = np.full(h * w, -1, dtype=np.int32)
    out[pix] = idx.astype(np.int32)
    return out


# ----------------------------------------------------------------------------------------------------------
# kernel (2): z-buffer splat
# ----------------------------------------------------------------------------------------------------------


def zbuffer_splat(points, K, w, h, quat, t, eye, lookat, face_normals, max_depth, depth_fmt=DEPTH_U16_MM):
    """Depth image of the cloud itself seen from one frame: per pixel the minimum over in-frustum points with
    that floor-pixel of the camera-space z (`uv[2]` of `camera_utils.py:23`, the row `points2pixel` discards),
    quantised as z_mm = floor(z*1000 + 0.5) clamped to [1, 65535] (uint16, 0 = hole) or kept as the float32
    rounding of the float64 minimum (DEPTH_F32_M).  No reference line produces this image (the reference reads
    sensor depth); it is the synthetic-scene generator of SURVEY 8(d)."""
    points = np.asarray(points, dtype=np.float64)
    ppts, pnrm = frame_planes(eye, lookat, face_normals, max_depth)
    idx = np.nonzero(point_inside_polyhedra(points, ppts, pnrm))[0]
    out = np.zeros(h * w, dtype=np.uint16 if depth_fmt == DEPTH_U16_MM else np.float32)
    if len(idx) == 0:
        return out.reshape(h, w)
    h0, h1, h2 = project_homogeneous(points[idx], K, quat, t)
    u, v = np.floor(h0 / h2), np.floor(h1 / h2)
    ok = (u >= 0) & (u < w) & (v >= 0) & (v < h)
    pix = (v[ok].astype(np.int64) * w + u[ok].astype(np.int64))
    z = h2[ok]
    if depth_fmt == DEPTH_U16_MM:
        q = np.clip(np.floor(z * 1000 + 0.5), 1, 65535).astype(np.int64)
        buf = np.full(h * w, 1 << 30, dtype=np.int64)
        np.minimum.at(buf, pix, q)
        out[:] = np.where(buf == (1 << 30), 0, buf).astype(np.uint16)
    else:
        zf = z.astype(np.float32)
        buf = np.full(h * w, np.inf, dtype=np.float32)
        np.minimum.at(buf, pix, zf)
        out[:] = np.where(np.isinf(buf), 0, buf)
    return out.reshape(h, w)


def zero_border(depth, border=10):
    """10-px zero border the reference multiplies into depth when `padding` is set (`ios_rtab.py:105-109`)."""
    d = depth.copy()
    if border > 0:                      # (a `-0:` slice would select the whole array)
        d[:border, :] = 0
        d[-border:, :] = 0
        d[:, :border] = 0
        d[:, -border:] = 0
    return d


# ----------------------------------------------------------------------------------------------------------
# level V: uv2pt + mask vote (a-10) and resize
# ----------------------------------------------------------------------------------------------------------


def resize_nearest(mask, w, h):
    """`cv2.resize(mask, (w, h), interpolation=cv2.INTER_NEAREST)` (`segUtils/voting.py:93`): OpenCV's rule
    sx = min(floor(dx * (src_w / dst_w)), src_w - 1) evaluated in float64 (same for rows)."""
    sh, sw = mask.shape
    fx = 1.0 / (w / sw)          # OpenCV: inv_scale_x = dsize.width / ssize.width ; fx = 1 / inv_scale_x
    fy = 1.0 / (h / sh)
    xs = np.minimum(np.floor(np.arange(w) * fx).astype(np.int64), sw - 1)
    ys = np.minimum(np.floor(np.arange(h) * fy).astype(np.int64), sh - 1)
    return mask[ys[:, None], xs[None, :]]


def vote_uv2pt(votes, uv2pt, mask):
    """One frame of `VotingSegmentation.vote` (`segUtils/voting.py:94-98`).  numpy's buffered fancy-index
    `+= 1` counts each distinct (point, class) pair of the frame ONCE."""
    mask = np.asarray(mask).reshape(-1)
    valid = uv2pt != -1
    if valid.any():
        votes[uv2pt[valid], mask[valid]] += 1
    return votes


# ----------------------------------------------------------------------------------------------------------
# kernel (3): label resolve (a-11)
# ----------------------------------------------------------------------------------------------------------


def segment(votes, nclasses, threshold=0.5, filter_classes=None):
    """`VotingSegmentation.segment` (`segUtils/voting.py:106-137`).  votes [N, C1] (any integer-valued
    dtype), `nclasses` = the id written for "unclassified" (self.nclasses).  -> int64 [N]."""
    votes = np.asarray(votes, dtype=np.float64)
    total = votes.sum(-1)                                        # voting.py:120
    sub = votes if filter_classes is None else votes[:, list(filter_classes)]   # voting.py:121
    valid = total > 0
    pc = np.argmax(sub, axis=1).astype(np.int64)                 # voting.py:124 (first maximum)
    pm = sub[np.arange(len(sub)), pc]
    pc[~valid] = nclasses                                        # voting.py:126
    prob = pm[valid] / total[valid]
    pc[np.where(valid)[0][prob < threshold]] = nclasses          # voting.py:128-130
    pc[pm == 0] = nclasses                                       # voting.py:131
    if filter_classes is not None:
        for i, c in enumerate(filter_classes):                   # voting.py:133-135 (sequential, aliasing)
            pc[pc == i] = c
    return pc


# ----------------------------------------------------------------------------------------------------------
# kernel (4): instance-box pair predicate, union-find closure and the sequential merge_bb driver (a-14/15)
# ----------------------------------------------------------------------------------------------------------


def aabb_overlap(lo_a, hi_a, lo_b, hi_b):
    """Closed-interval per-axis overlap of `check_intersection` (`merge_intersecting_bb.py:51-53`):
    (min1<=min2<=max1) or (min2<=min1<=max2) on x, y and z."""
    ok = True
    for k in range(3):
        ok = ok and ((lo_a[k] <= lo_b[k] <= hi_a[k]) or (lo_b[k] <= lo_a[k] <= hi_b[k]))
    return ok


def box_pairs_aabb(lo, hi, group):
    """All unordered pairs (i<j) with equal `group` (category / parent id, `merge_intersecting_bb.py:49,80`)
    whose AABBs overlap.  Sort-and-sweep on x so it is usable at 200 k boxes.  -> int64 [E,2] sorted."""
    lo = np.asarray(lo, dtype=np.float64)
    hi = np.asarray(hi, dtype=np.float64)
    group = np.asarray(group)
    order = np.argsort(lo[:, 0], kind="stable")
    slo = lo[order, 0]
    edges = []
    for a_pos, a in enumerate(order):
        # candidates b with lo_x[b] in [lo_x[a], hi_x[a]]  (then (min1<=min2<=max1) holds on x)
        end = np.searchsorted(slo, hi[a, 0], side="right")
        cand = order[a_pos + 1:end]
        if len(cand) == 0:
            continue
        m = group[cand] == group[a]
        for k in (1, 2):
            m &= ((lo[a, k] <= lo[cand, k]) & (lo[cand, k] <= hi[a, k])) | \
                 ((lo[cand, k] <= lo[a, k]) & (lo[a, k] <= hi[cand, k]))
        for b in cand[m]:
            edges.append((min(a, b), max(a, b)))
    if not edges:
        return np.zeros((0, 2), dtype=np.int64)
    e = np.unique(np.asarray(edges, dtype=np.int64), axis=0)
    return e


def union_find_labels(n, edges):
    """Transitive closure of the pair predicate: label = smallest box index in the connected component."""
    parent = np.arange(n, dtype=np.int64)

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    for a, b in np.asarray(edges, dtype=np.int64).reshape(-1, 2):
        ra, rb = find(a), find(b)
        if ra != rb:
            if ra < rb:
                parent[rb] = ra
            else:
                parent[ra] = rb
    return np.array([find(i) for i in range(n)], dtype=np.int64)


def obb_contains(center, R, extent, pts):
    """Open3D `OrientedBoundingBox.get_point_indices_within_bounding_box` membership rule
    (call sites `merge_intersecting_bb.py:76,87`): |(p-c).R[:,k]| <= extent[k]/2 for k = 0..2 (closed)."""
    d = np.asarray(pts, dtype=np.float64) - np.asarray(center, dtype=np.float64)[None, :]
    R = np.asarray(R, dtype=np.float64)
    ok = np.ones(len(d), dtype=bool)
    for k in range(3):
        proj = (d[:, 0] * R[0, k] + d[:, 1] * R[1, k]) + d[:, 2] * R[2, k]
        ok &= np.abs(proj) <= extent[k] / 2
    return ok


def fit_box(points, model="pca"):
    """The STATED box models behind the `o3d.geometry.OrientedBoundingBox.create_from_points` call sites
    (`merge_intersecting_bb.py:18,75,86,126`, `get3DSeg.py:434`) -- identical to `oracle/refshim/open3d`, which is what
    the unmodified reference was run on to make `tests/golden/g7_merge.json`.  Open3D's own fit (Qhull hull vertices ->
    covariance) is NOT reproduced: parity of the FIT is unpinned, parity of everything downstream of a box is pinned.
    -> (centre [3], R [3,3] columns = axes, extent [3])."""
    p = np.asarray(points, dtype=np.float64)
    if model == "aabb":
        mn, mx = p.min(0), p.max(0)
        return (mn + mx) * 0.5, np.eye(3), mx - mn
    mean = p.mean(0)
    q = p - mean
    cov = (q.T @ q) / max(len(p) - 1, 1)
    _, evecs = np.linalg.eigh(cov)
    R = evecs[:, [2, 1, 0]].copy()
    R[:, 2] = np.cross(R[:, 0], R[:, 1])
    proj = q @ R
    mn, mx = proj.min(0), proj.max(0)
    return mean + R @ ((mn + mx) * 0.5), R, mx - mn


def box_corners(center, R, extent):
    """`OrientedBoundingBox.get_box_points` (call sites `merge_intersecting_bb.py:19,127`, `get3DSeg.py:435`):
    the 8 corners centre +- R[:,k] * extent[k] / 2 in Open3D's corner order."""
    x, y, z = (R[:, k] * (extent[k] * 0.5) for k in range(3))
    c = np.asarray(center, dtype=np.float64)
    return np.array([c - x - y - z, c + x - y - z, c - x + y - z, c - x - y + z, c + x + y + z, c - x + y + z,
                     c + x - y + z, c + x + y - z])


def cal_min_max(corners):
    """`cal_min_max` (`merge_intersecting_bb.py:15-42`) from the 8 box corners: each corner is projected on the x, y
    and z axis through the origin (`Line.project_point`, `:20-36`), the component-wise min / max of the projected
    points is taken and its EXACT-ZERO components are dropped (`min_x[np.nonzero(min_x)]`, `:23-25`) -- so every
    result is a 1-element array (empty if the bound is exactly 0).  -> (min_x, max_x, min_y, max_y, min_z, max_z)."""
    c = np.asarray(corners, dtype=np.float64)
    out = []
    for k in range(3):
        proj = np.zeros_like(c)
        proj[:, k] = c[:, k]                                     # point + direction * ((p - point) . direction) / 1
        for v in (proj.min(0), proj.max(0)):
            out.append(v[np.nonzero(v)])
    return tuple(out)


def check_intersection_as_shipped(id1, id_list, ids, pts, info_sem, model="pca"):
    """`check_intersection` (`merge_intersecting_bb.py:44-56`) as shipped: the result list is re-created inside the
    loop (`:48`), so only the LAST id2 != id1 decides what is returned; same-category gate `:49`; closed-interval
    AABB-of-corners overlap `:51-53`."""
    def mm(k):
        sel = ids == k
        return cal_min_max(box_corners(*fit_box(pts[sel], model)))
    a = mm(id_list[id1])
    intersecting = []
    for id2 in range(1, len(id_list)):
        if id1 != id2:
            intersecting = []
            if info_sem[id1]["category_id"] == info_sem[id2]["category_id"]:
                b = mm(id_list[id2])
                if all(((a[2 * k] <= b[2 * k]) and (b[2 * k] <= a[2 * k + 1])) or ((b[2 * k] <= a[2 * k]) and (a[2 * k] <= b[2 * k + 1]))
                       for k in range(3)):
                    intersecting.append(id2)
    return intersecting


def merge_hit_fn(pts, model="pca"):
    """`hit_fn` for merge_bb_sequential: the geometric part of `check_intersection_open3d` (`:68-91`) -- box of the
    instance's points, membership of the WHOLE cloud (`:75-76,86-87`), non-empty intersection of the index lists (`:88-90`)."""
    pts = np.asarray(pts, dtype=np.float64)

    def inside(k, ids):
        sel = ids == k
        if sel.sum() < 4:
            return None
        return obb_contains(*fit_box(pts[sel], model), pts)

    def hit(id1, id2, ids):
        a = inside(id1, ids)
        if id2 is None:
            return False if a is None else True
        b = inside(id2, ids)
        if b is None:
            return None
        return bool((a & b).any())
    return hit


def merge_bb_final_boxes(info_sem, ids, pts, model="pca"):
    """Tail of `merge_bb` (`merge_intersecting_bb.py:122-128`): fresh box corners for survivors with > 4 points."""
    for k in range(1, len(info_sem)):
        sel = ids == info_sem[k]["id"]
        if sel.sum() > 4:
            info_sem[k]["bbox"] = box_corners(*fit_box(pts[sel], model)).tolist()
    return info_sem


def merge_bb_sequential(info_sem, ids, hit_fn):
    """The order-dependent driver of `merge_bb` (`merge_intersecting_bb.py:103-120`) with its quirks kept:
    loop index used as instance id (`:70,113`), `< len(info_sem)-1` guards on the shrinking list (`:79`),
    early `return` when an id2 has < 4 points (`:83-84`), `del info_sem[i]` without index correction
    (`:118-120`).  `hit_fn(id1, id2, ids)` -> None if id2 has < 4 points else bool (boxes share a cloud point);
    `hit_fn(id1, None, ids)` -> False if id1 has < 4 points.  Mutates and returns (info_sem, ids)."""
    L = len(info_sem)
    for id1 in range(1, L):
        hits = []
        if hit_fn(id1, None, ids) is not False:
            for id2 in range(1, L):
                if id1 != id2 and id2 < len(info_sem) - 1 and id1 < len(info_sem) - 1:
                    if info_sem[id1]["parent_id"] == info_sem[id2]["parent_id"]:
                        r = hit_fn(id1, id2, ids)
                        if r is None:
                            break                                 # early return, :83-84
                        if r:
                            hits.append(id2)
        if hits:
            for hb in hits:                                       # update_id_info, :58-62
                sel = ids == hb
                info_sem[id1]["area"] += info_sem[hb]["area"]
                ids[sel] = id1
            for i in hits:                                        # :118-120
                if i < len(info_sem):
                    del info_sem[i]
    return info_sem, ids


# ----------------------------------------------------------------------------------------------------------
# next row: instance split
# ----------------------------------------------------------------------------------------------------------


def split_into_instances(classes, indptr, indices, nclasses=133, instance_classes=None, minimum_points=1):
    """`split_into_instances` (`Fusion3DSeg/segUtils/cv.py:402-500`) restated on a CSR adjacency (indptr, indices):
    connected components of equal class, numbered class by class in `instance_classes` order and, inside a class,
    by ascending smallest point index (the BFS seeds are `remaining_points[0]`, `cv.py:473-475`); components with
    fewer than `minimum_points` points are re-classed `nclasses` and folded into one "small disjoint" instance
    (`cv.py:478-486`).  Returns (ninstances, ids [N], info list, classes [N])."""
    classes = np.asarray(classes).copy()
    n = len(classes)
    allclasses = np.unique(classes)
    ids = np.zeros_like(classes)
    info = []
    small_id = None
    if instance_classes is None:                                   # cv.py:448-456
        inst = allclasses
        ninst, sem = 0, []
        if (inst == nclasses).any():
            inst = inst[inst != nclasses]
            sem, ninst = [nclasses], 1
    else:                                                          # cv.py:457-460
        inst = np.array(instance_classes)
        sem = np.setdiff1d(allclasses, inst)
        ninst = len(sem)
    for k in range(ninst if len(sem) else 0):                      # cv.py:462-470
        m = classes == sem[k]
        ids[m] = k
        info.append({'id': k, 'isthing': False, 'category_id': int(sem[k]), 'area': int(m.sum())})
        if sem[k] == nclasses:
            small_id = k
    for c in inst:                                                 # cv.py:472-499
        todo = classes == c
        for seed in np.nonzero(todo)[0]:
            if not todo[seed]:
                continue
            comp, stack, seen = [], [seed], {int(seed)}
            while stack:                                           # flood fill over equal-class neighbours (cv.py:425-440)
                p = stack.pop()
                comp.append(p)
                for q in indices[indptr[p]:indptr[p + 1]]:
                    q = int(q)
                    if q not in seen and classes[q] == c and todo[q]:
                        seen.add(q)
                        stack.append(q)
            comp = np.array(comp)
            if len(comp) < minimum_points:
                if small_id is None:
                    small_id = ninst
                    info.append({'id': ninst, 'isthing': True, 'category_id': int(nclasses), 'area': 0})
                    ninst += 1
                info[small_id]['area'] += int(len(comp))
                ids[comp] = small_id
                newc = nclasses
            else:
                info.append({'id': ninst, 'isthing': True, 'category_id': int(c), 'area': int(len(comp))})
                ids[comp] = ninst
                ninst += 1
                newc = c
            todo[comp] = False
            classes[comp] = newc
    return ninst, ids, info, classes


# ----------------------------------------------------------------------------------------------------------
# radius adjacency: `KDTree(points).query_radius(points, r=2*ds_radius)`, fusion.py:374-375
# ----------------------------------------------------------------------------------------------------------


def radius_adjacency(points, r, chunk=2048):
    """CSR (indptr int64 [N+1], indices int64) of `sklearn.neighbors.KDTree(points).query_radius(points, r)` with every
    row sorted ascending (scikit-learn returns a row in tree-traversal order; `split_into_instances` only uses the
    rows as sets, `segUtils/cv.py:425-440`).  Membership is scikit-learn's leaf test for the Euclidean metric: the
    reduced distance `((dx*dx) + dy*dy) + dz*dz` (binary64, left to right, never fused) `<= r*r`; a point is its own
    neighbour.  Restated on a uniform grid of cell size r (27 cells per query) so that it finishes on 1e5-point clouds.
    Pinned by `tests/test_oracle_golden.py::test_radius_adjacency_matches_kdtree` against scikit-learn itself."""
    p = np.ascontiguousarray(np.asarray(points, dtype=np.float64))
    n = len(p)
    r = float(r)
    r2 = r * r
    if n == 0:
        return np.zeros(1, np.int64), np.zeros(0, np.int64)
    cell = np.floor((p - p.min(axis=0)) / max(r, 1e-300)).astype(np.int64)
    dims = cell.max(axis=0) + 3
    key = ((cell[:, 0] + 1) * dims[1] + (cell[:, 1] + 1)) * dims[2] + (cell[:, 2] + 1)
    order = np.argsort(key, kind="stable")
    skey = key[order]
    rows = [None] * n
    offs = [((dx * dims[1]) + dy) * dims[2] + dz for dx in (-1, 0, 1) for dy in (-1, 0, 1) for dz in (-1, 0, 1)]
    uniq, start = np.unique(skey, return_index=True)
    end = np.append(start[1:], n)
    for u, a, b in zip(uniq, start, end):
        q = order[a:b]                                    # the points of this cell
        lo = np.searchsorted(skey, [u + o for o in offs], side="left")
        hi = np.searchsorted(skey, [u + o for o in offs], side="right")
        cand = np.concatenate([order[x:y] for x, y in zip(lo, hi) if y > x])
        for c0 in range(0, len(q), chunk):
            qq = q[c0:c0 + chunk]
            dx = p[qq, None, 0] - p[None, cand, 0]
            dy = p[qq, None, 1] - p[None, cand, 1]
            dz = p[qq, None, 2] - p[None, cand, 2]
            d2 = (dx * dx + dy * dy) + dz * dz
            hit = d2 <= r2
            for k, i in enumerate(qq):
                rows[i] = np.sort(cand[hit[k]])
    indptr = np.zeros(n + 1, np.int64)
    indptr[1:] = np.cumsum([len(x) for x in rows])
    return indptr, np.concatenate(rows).astype(np.int64)
