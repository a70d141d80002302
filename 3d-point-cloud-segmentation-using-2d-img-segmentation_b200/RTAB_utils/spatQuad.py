"""The quaternion semantics the reference gets from `RTAB_utils/spatQuad.py` + pyquaternion, as plain data:
only construction / inverse are host-side (pose bookkeeping); the rotation itself runs inside the CUDA kernels
(`dquat_rotate` in csrc/f3d_common.cuh restates `SpatQuadranion.rotate`, spatQuad.py:16-28)."""
from __future__ import annotations

import numpy as np


def wxyz_from_pose_text(xyzw):
    """`parse_rts` re-order (Fusion3DSeg/fusion.py:71-72): pose text stores (x, y, z, w); the path uses (w, x, y, z)."""
    a = np.asarray(xyzw, dtype=np.float64)
    return np.ascontiguousarray(a[..., [3, 0, 1, 2]])
