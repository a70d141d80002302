"""Seeded synthetic scenes of the shapes BASELINE.json names.

Pure numpy, host side, no reference code involved: a room/building cloud rounded to float32 and Morton
sorted, an RTAB-style closed-loop trajectory whose quaternions are rounded to 6 decimals in (x, y, z, w)
text order and re-ordered to (w, x, y, z) exactly like `parse_rts` does (`Fusion3DSeg/fusion.py:71-72`),
and intrinsics = `RTAB_utils/calibration.yaml:6-10` scaled as `RTAB_utils/ios_rtab.py:125-131`.
Depth images are NOT produced here: they are the z-buffer splat of the cloud (CUDA kernel 2 in the product,
`oracle.zbuffer_splat` in tests).  Block-constant label masks are produced by `block_masks`.
"""
from __future__ import annotations

import dataclasses

import numpy as np

# RTAB_utils/calibration.yaml:4-10  (720 x 960 RGB camera)
CALIB_W, CALIB_H = 720, 960
CALIB_FX = 7.9894403076171875e+02
CALIB_FY = 7.9894403076171875e+02
CALIB_CX = 3.6195578002929688e+02
CALIB_CY = 4.7456329345703125e+02


@dataclasses.dataclass
class SceneSpec:
    name: str
    npoints: int
    nframes: int
    width: int
    height: int
    room: tuple          # (Lx, Ly, Lz) metres per storey
    storeys: int
    nobjects: int
    zmax: float          # valid depth range upper bound == far plane distance (process3D.py:17,39)
    seed: int


CONFIGS = {
    # BASELINE.json configs[0..3]; seeds 1000 + config index
    "C1": SceneSpec("C1", 1_000_000, 50, 640, 480, (8.0, 6.0, 3.0), 1, 40, 4.0, 1000),
    "C2": SceneSpec("C2", 10_000_000, 500, 1920, 1440, (20.0, 15.0, 3.0), 1, 200, 4.0, 1001),
    "C3": SceneSpec("C3", 100_000_000, 5000, 1920, 1440, (60.0, 40.0, 3.0), 20, 10_000, 10.0, 1002),
    "C4": SceneSpec("C4", 20_000_000, 1000, 3840, 2160, (20.0, 15.0, 3.0), 1, 200, 10.0, 1003),
}


def scaled_spec(base: str, npoints=None, nframes=None, width=None, height=None, seed=None) -> SceneSpec:
    s = CONFIGS[base]
    return dataclasses.replace(
        s, npoints=npoints or s.npoints, nframes=nframes or s.nframes, width=width or s.width,
        height=height or s.height, seed=s.seed if seed is None else seed)


def scaled_intrinsics(width: int, height: int) -> np.ndarray:
    """`__resize_camera_matrix(Depth_W/RGB_W, Depth_H/RGB_H)` (`RTAB_utils/ios_rtab.py:115-131,167`)."""
    sx, sy = width / CALIB_W, height / CALIB_H
    return np.array([[CALIB_FX * sx, 0.0, CALIB_CX * sx], [0.0, CALIB_FY * sy, CALIB_CY * sy], [0.0, 0.0, 1.0]])


def _sample_box_faces(rng, n, lo, hi, faces):
    """n points uniform by area on the listed faces (axis, side) of the axis-aligned box [lo, hi]."""
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    ext = hi - lo
    areas = np.array([ext[(a + 1) % 3] * ext[(a + 2) % 3] for a, _ in faces])
    counts = rng.multinomial(n, areas / areas.sum())
    out = []
    for (a, side), c in zip(faces, counts):
        p = lo[None, :] + rng.random((c, 3)) * ext[None, :]
        p[:, a] = hi[a] if side else lo[a]
        out.append(p)
    return np.concatenate(out, axis=0)


def morton_order(pts: np.ndarray) -> np.ndarray:
    """Permutation sorting float points by 3-D Morton code (21 bits per axis)."""
    lo = pts.min(0)
    ext = np.maximum(pts.max(0) - lo, 1e-9)
    q = np.minimum(((pts - lo) / ext * (1 << 21)).astype(np.uint64), (1 << 21) - 1)

    def spread(x):
        x = x & np.uint64(0x1FFFFF)
        x = (x | (x << np.uint64(32))) & np.uint64(0x1F00000000FFFF)
        x = (x | (x << np.uint64(16))) & np.uint64(0x1F0000FF0000FF)
        x = (x | (x << np.uint64(8))) & np.uint64(0x100F00F00F00F00F)
        x = (x | (x << np.uint64(4))) & np.uint64(0x10C30C30C30C30C3)
        x = (x | (x << np.uint64(2))) & np.uint64(0x1249249249249249)
        return x

    code = spread(q[:, 0]) | (spread(q[:, 1]) << np.uint64(1)) | (spread(q[:, 2]) << np.uint64(2))
    return np.argsort(code, kind="stable")


def make_cloud(spec: SceneSpec) -> np.ndarray:
    """float32 [N,3] room cloud: inner faces of `storeys` stacked boxes + cuboids standing on the floors,
    1 mm Gaussian jitter, Morton sorted."""
    rng = np.random.Generator(np.random.PCG64(spec.seed))
    Lx, Ly, Lz = spec.room
    n_obj_pts = int(spec.npoints * 0.3) if spec.nobjects else 0
    n_room = spec.npoints - n_obj_pts
    all_faces = [(0, 0), (0, 1), (1, 0), (1, 1), (2, 0), (2, 1)]
    parts = []
    per_storey = np.full(spec.storeys, n_room // spec.storeys)
    per_storey[: n_room % spec.storeys] += 1
    for s in range(spec.storeys):
        z0 = s * Lz
        parts.append(_sample_box_faces(rng, int(per_storey[s]), (0, 0, z0), (Lx, Ly, z0 + Lz), all_faces))
    if spec.nobjects:
        sizes = np.stack([rng.uniform(0.3, 1.5, spec.nobjects), rng.uniform(0.3, 1.5, spec.nobjects),
                          rng.uniform(0.3, 2.0, spec.nobjects)], axis=1)
        origin = np.stack([rng.uniform(0.2, Lx - 1.7, spec.nobjects), rng.uniform(0.2, Ly - 1.7, spec.nobjects),
                           rng.integers(0, spec.storeys, spec.nobjects) * Lz], axis=1)
        area = 2 * sizes[:, 2] * (sizes[:, 0] + sizes[:, 1]) + sizes[:, 0] * sizes[:, 1]
        counts = rng.multinomial(n_obj_pts, area / area.sum())
        faces5 = [(0, 0), (0, 1), (1, 0), (1, 1), (2, 1)]
        for o in range(spec.nobjects):
            if counts[o]:
                parts.append(_sample_box_faces(rng, int(counts[o]), origin[o], origin[o] + sizes[o], faces5))
    pts = np.concatenate(parts, axis=0)
    pts += rng.normal(0.0, 1e-3, pts.shape)
    pts = pts.astype(np.float32)
    return np.ascontiguousarray(pts[morton_order(pts)])


def _rot_to_quat_wxyz(R):
    """Rotation matrix (camera->world) to unit quaternion (w, x, y, z)."""
    tr = R[0, 0] + R[1, 1] + R[2, 2]
    if tr > 0:
        s = np.sqrt(tr + 1.0) * 2
        return np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    i = int(np.argmax([R[0, 0], R[1, 1], R[2, 2]]))
    j, k = (i + 1) % 3, (i + 2) % 3
    s = np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
    q = np.zeros(4)
    q[0] = (R[k, j] - R[j, k]) / s
    q[1 + i] = 0.25 * s
    q[1 + j] = (R[j, i] + R[i, j]) / s
    q[1 + k] = (R[k, i] + R[i, k]) / s
    return q


def make_poses(spec: SceneSpec):
    """Closed-loop trajectory at 1.5 m above each floor.  Returns (wxyz [F,4], t [F,3]) float64, both rounded
    to 6 decimals as the `rtabmap-export` pose text is (quaternions are therefore NOT unit; a-2)."""
    rng = np.random.Generator(np.random.PCG64(spec.seed + 7919))
    Lx, Ly, Lz = spec.room
    F = spec.nframes
    s = np.arange(F) / F
    storey = np.minimum((s * spec.storeys).astype(int), spec.storeys - 1)
    ang = 2 * np.pi * s * max(1, spec.storeys) * 1.0
    cx, cy = Lx / 2, Ly / 2
    pos = np.stack([cx + 0.3 * Lx * np.cos(ang), cy + 0.3 * Ly * np.sin(ang), storey * Lz + 1.5], axis=1)
    tang = np.stack([-0.3 * Lx * np.sin(ang), 0.3 * Ly * np.cos(ang)], axis=1)
    yaw = np.arctan2(tang[:, 1], tang[:, 0]) + np.deg2rad(rng.normal(0, 10, F))
    pitch = np.deg2rad(rng.normal(0, 10, F))
    roll = np.deg2rad(rng.normal(0, 2, F))
    quats = np.zeros((F, 4))
    for f in range(F):
        fwd = np.array([np.cos(yaw[f]) * np.cos(pitch[f]), np.sin(yaw[f]) * np.cos(pitch[f]), np.sin(pitch[f])])
        up = np.array([0.0, 0.0, 1.0])
        right = np.cross(fwd, up)
        right /= np.linalg.norm(right)
        down = np.cross(fwd, right)
        cr, sr = np.cos(roll[f]), np.sin(roll[f])
        right, down = cr * right + sr * down, -sr * right + cr * down
        R = np.stack([right, down, fwd], axis=1)          # columns: camera x (right), y (down), z (forward)
        quats[f] = _rot_to_quat_wxyz(R)
    xyzw = np.round(quats[:, [1, 2, 3, 0]], 6)             # pose text order (ios_rtab.py:61-68)
    wxyz = xyzw[:, [3, 0, 1, 2]]                           # parse_rts re-order (fusion.py:71-72)
    return np.ascontiguousarray(wxyz), np.round(pos, 6)


def block_masks(spec_or_shape, nframes=None, seed=0, nclasses=133, block=32, unclassified_frac=0.05):
    """uint8 [F,H,W] label images, piecewise constant on `block` x `block` pixels, labels 0..nclasses-1 with
    `unclassified_frac` of blocks = nclasses (133 = unclassified, `get2DSeg.py:118`)."""
    if isinstance(spec_or_shape, SceneSpec):
        F, H, W, seed = spec_or_shape.nframes, spec_or_shape.height, spec_or_shape.width, spec_or_shape.seed
    else:
        H, W = spec_or_shape
        F = nframes
    rng = np.random.Generator(np.random.PCG64(seed + 104729))
    bh, bw = -(-H // block), -(-W // block)
    lab = rng.integers(0, nclasses, (F, bh, bw), dtype=np.int64)
    lab[rng.random((F, bh, bw)) < unclassified_frac] = nclasses
    m = np.repeat(np.repeat(lab, block, axis=1), block, axis=2)[:, :H, :W]
    return np.ascontiguousarray(m.astype(np.uint8))


def make_boxes(nboxes=200_000, seed=1004, extent=(80.0, 80.0, 5.0), ngroups=16):
    """C5: axis-aligned instance boxes, dense enough that a box overlaps ~2-3 others of its group.  Returns lo [B,3], hi [B,3] float64, group int32 [B], area int64 [B]."""
    rng = np.random.Generator(np.random.PCG64(seed))
    c = rng.random((nboxes, 3)) * np.asarray(extent)[None, :]
    half = np.exp(rng.normal(np.log(0.4), 0.5, (nboxes, 3)))
    group = rng.integers(0, ngroups, nboxes).astype(np.int32)
    area = rng.integers(4, 5001, nboxes).astype(np.int64)
    return c - half, c + half, group, area
