"""Drop-in for `RTAB_utils/spatQuad.py`: `SpatQuadranion` with the pyquaternion semantics the reference relies on
(construction from (w, x, y, z) WITHOUT normalisation, `.elements`, `.inverse` = conjugate / sum of squares, `.w/.x/.y/.z`;
third-party `pyquaternion` is neither vendored by the reference nor needed here) and `rotate` on the GPU
(`f3d_quat_rotate`: the raw Hamilton sandwich of `spatQuad.py:16-28` in float64, reference operation order)."""
from __future__ import annotations

import numpy as np

from .. import engine


class SpatQuadranion:
    def __init__(self, *args, array=None):
        if array is not None:
            q = np.asarray(array, dtype=np.float64)
        elif len(args) == 1:
            q = np.asarray([float(a) for a in args[0]], dtype=np.float64)      # strings allowed, as `ios_rtab.py:188` passes them
        elif len(args) == 4:
            q = np.asarray([float(a) for a in args], dtype=np.float64)
        else:
            raise ValueError("SpatQuadranion(w, x, y, z) / SpatQuadranion(seq4) / SpatQuadranion(array=)")
        if q.shape != (4,):
            raise ValueError("a quaternion needs 4 elements")
        self.q = q

    elements = property(lambda self: self.q)
    w = property(lambda self: self.q[0])
    x = property(lambda self: self.q[1])
    y = property(lambda self: self.q[2])
    z = property(lambda self: self.q[3])

    @property
    def inverse(self):
        """pyquaternion `Quaternion.inverse`: conjugate / sum of squares ((w*w + x*x) + y*y) + z*z, no normalisation."""
        q = self.q
        ss = ((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3]
        if not ss > 0:
            raise ZeroDivisionError("a zero quaternion cannot be inverted")
        return SpatQuadranion(array=np.array([q[0] / ss, -q[1] / ss, -q[2] / ss, -q[3] / ss]))

    def rotate(self, p):
        """[N,3] -> [N,3] float64 (`spatQuad.py:6-28`)."""
        p = np.ascontiguousarray(np.asarray(p, dtype=np.float64).reshape(-1, 3))
        return engine.quat_rotate(p, self.q).cpu().numpy()

    def __str__(self):
        return f'TooliqaQuaternion{self.w, self.x, self.y, self.z}'

    __repr__ = __str__


def wxyz_from_pose_text(xyzw):
    """`parse_rts` re-order (Fusion3DSeg/fusion.py:71-72): pose text stores (x, y, z, w); the path uses (w, x, y, z)."""
    a = np.asarray(xyzw, dtype=np.float64)
    return np.ascontiguousarray(a[..., [3, 0, 1, 2]])
