"""Drop-in for the one predicate of `Fusion3DSeg/intersections.py` that is on the label-fusion path."""
from __future__ import annotations

import numpy as np

from .. import engine


def point_inside_polyhedra(points, plane_points, normals):
    """Same contract as the reference (`Fusion3DSeg/intersections.py:146-164`): inside <=> (p - a_m).n_m >= 0 for
    every plane m.  points [N,3], plane_points [M,3], normals [M,3] (inward) -> bool [N].  GPU float64
    (`f3d_frustum_mask`)."""
    pts = np.ascontiguousarray(np.asarray(points, dtype=np.float64).reshape(-1, 3))
    m = engine.frustum_mask(pts, plane_points, normals)
    return m.cpu().numpy().astype(bool)
