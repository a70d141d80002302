"""TEST INFRASTRUCTURE ONLY -- minimal stand-in for the third-party `pyquaternion` package.

The reference (`/root/reference/RTAB_utils/spatQuad.py:2,6`) subclasses `pyquaternion.Quaternion`, which is
not installed in this image (no network).  Only the semantics the reference's hot path touches are provided:
construction from a 4-sequence / 4 scalars (strings allowed, `ios_rtab.py:188`) in (w, x, y, z) order WITHOUT
normalisation, `.elements`, `.inverse` (= conjugate / sum of squares) and `.w/.x/.y/.z`.
It is used exclusively by `tests/golden/make_golden.py` to run the unmodified reference in the build
container.  Nothing in the product imports it.
"""
import numpy as np


class Quaternion:
    def __init__(self, *args, **kwargs):
        if "array" in kwargs:
            q = np.asarray(kwargs["array"], dtype=np.float64)
        elif len(args) == 1:
            q = np.asarray([float(a) for a in args[0]], dtype=np.float64)
        elif len(args) == 4:
            q = np.asarray([float(a) for a in args], dtype=np.float64)
        else:
            raise ValueError("shim supports Quaternion(seq4) / Quaternion(w,x,y,z) / Quaternion(array=)")
        if q.shape != (4,):
            raise ValueError("quaternion needs 4 elements")
        self.q = q

    @property
    def elements(self):
        return self.q

    def _sum_of_squares(self):
        return np.dot(self.q, self.q)

    def _vector_conjugate(self):
        return np.hstack((self.q[0], -self.q[1:4]))

    @property
    def inverse(self):
        ss = self._sum_of_squares()
        if ss > 0:
            return self.__class__(array=(self._vector_conjugate() / ss))
        raise ZeroDivisionError("a zero quaternion cannot be inverted")

    @property
    def w(self):
        return self.q[0]

    @property
    def x(self):
        return self.q[1]

    @property
    def y(self):
        return self.q[2]

    @property
    def z(self):
        return self.q[3]
