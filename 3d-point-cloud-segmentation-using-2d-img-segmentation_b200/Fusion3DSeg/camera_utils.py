"""Drop-in for `Fusion3DSeg/camera_utils.py` of the reference: world -> pixel projection on the GPU."""
from __future__ import annotations

import numpy as np

from .. import engine


def points2pixel(points, intrinsic, quat, translation):
    """Same contract as the reference `points2pixel` (`Fusion3DSeg/camera_utils.py:9-26`).

    Args:
        points (np.ndarray[float]): [N, 3] xyz points.
        intrinsic (np.ndarray[float]): [3, 3] intrinsic matrix.
        quat (np.ndarray[float]): [4] (w, x, y, z), NOT normalised (pyquaternion inverse semantics).
        translation (np.ndarray): [3] camera translation.

    Returns:
        np.ndarray[int32]: [2, N] floor pixel coordinates (row 0 = u, row 1 = v), evaluated in float64 on the GPU
        (`f3d_project_pixels`).
    """
    pts = np.ascontiguousarray(np.asarray(points, dtype=np.float64).reshape(-1, 3))
    uv = engine.project_pixels(pts, intrinsic, quat, translation)
    return uv.cpu().numpy()
