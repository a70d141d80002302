"""TEST INFRASTRUCTURE ONLY -- stand-in so that `import open3d as o3d` at the top of the reference's
`Fusion3DSeg/fusion.py:6` / `Fusion3DSeg/merge_intersecting_bb.py:9` succeeds in the build container (Open3D is not
installed here and not in the wheelhouse).

For `fusion.py` the module is only used for PLY I/O and GUI windows, never for arithmetic on the label-fusion path.
For `merge_intersecting_bb.py` the three `OrientedBoundingBox` calls the reference makes ARE arithmetic, so they are
backed here by an explicitly STATED box model -- not Open3D's (its `create_from_points` runs Qhull first and fits the
covariance of the hull vertices; that cannot be reproduced without Open3D):

    BOX_MODEL = "pca"   centre / axes from the mean and covariance of ALL points (numpy `eigh`, axes by descending
                        eigenvalue, third axis = first x second), extents = range of the projections, centre = mean +
                        R @ mid-range -- the box `Fusion3DSeg/merge_intersecting_bb.fit_obb` of the product fits;
    BOX_MODEL = "aabb"  axis-aligned box of the points (R = I) -- the stand-in SURVEY a-14's probe used.

`get_point_indices_within_bounding_box` is Open3D's published rule |(p - c) . R[:, k]| <= extent[k] / 2 (closed) and
`get_box_points` returns the 8 corners centre + R @ (+-e/2) in Open3D's corner order.  What the golden vectors made with
this shim pin is therefore the reference's DRIVER logic (`merge_bb`, `check_intersection_open3d`, `update_id_info`,
`cal_min_max`, `check_intersection`) on a stated box model, not Open3D's hull fit.
"""
import types

import numpy as np

BOX_MODEL = "pca"


class _Box:
    def __init__(self, center, R, extent):
        self.center, self.R, self.extent = center, R, extent
        self.color = None

    @staticmethod
    def create_from_points(points):
        p = np.asarray(points, dtype=np.float64)
        if BOX_MODEL == "aabb":
            mn, mx = p.min(0), p.max(0)
            return _Box((mn + mx) * 0.5, np.eye(3), mx - mn)
        mean = p.mean(0)
        q = p - mean
        cov = (q.T @ q) / max(len(p) - 1, 1)
        evals, evecs = np.linalg.eigh(cov)
        R = evecs[:, [2, 1, 0]].copy()
        R[:, 2] = np.cross(R[:, 0], R[:, 1])
        proj = q @ R
        mn, mx = proj.min(0), proj.max(0)
        return _Box(mean + R @ ((mn + mx) * 0.5), R, mx - mn)

    def get_point_indices_within_bounding_box(self, points):
        d = np.asarray(points, dtype=np.float64) - self.center[None, :]
        ok = np.ones(len(d), dtype=bool)
        for k in range(3):
            proj = (d[:, 0] * self.R[0, k] + d[:, 1] * self.R[1, k]) + d[:, 2] * self.R[2, k]
            ok &= np.abs(proj) <= self.extent[k] / 2
        return [int(i) for i in np.nonzero(ok)[0]]

    def get_box_points(self):
        # Open3D corner order (OrientedBoundingBox::GetBoxPoints)
        x, y, z = (self.R[:, k] * (self.extent[k] * 0.5) for k in range(3))
        c = self.center
        return np.array([c - x - y - z, c + x - y - z, c - x + y - z, c - x - y + z, c + x + y + z, c - x + y + z,
                         c + x - y + z, c + x + y - z])


class PointCloud:
    def __init__(self, points=None):
        self.points = np.zeros((0, 3)) if points is None else np.asarray(points, dtype=np.float64)


geometry = types.SimpleNamespace(OrientedBoundingBox=_Box, PointCloud=PointCloud)
utility = types.SimpleNamespace(Vector3dVector=lambda a: np.asarray(a, dtype=np.float64))


# ---- file I/O and GUI calls of `get3DSeg.master_classes` (`get3DSeg.py:379,465-466`): a binary little-endian PLY reader / writer for
# x y z (double) [+ uchar colours], and a no-op window -----------------------------------------------------------------------------
def _read_point_cloud(path):
    with open(path, "rb") as fp:
        props, n = [], 0
        while True:
            line = fp.readline().decode("ascii").strip()
            if line.startswith("element vertex"):
                n = int(line.split()[-1])
            elif line.startswith("property"):
                _, typ, name = line.split()
                props.append((name, {"double": "<f8", "float": "<f4", "uchar": "u1"}[typ]))
            elif line == "end_header":
                break
        rec = np.frombuffer(fp.read(), dtype=props, count=n)
    return PointCloud(np.stack([rec["x"], rec["y"], rec["z"]], axis=1))


def _write_point_cloud(path, pcd):
    pts = np.asarray(pcd.points, dtype=np.float64)
    cols = getattr(pcd, "colors", None)
    fields = [("x", "<f8"), ("y", "<f8"), ("z", "<f8")]
    header = ["ply", "format binary_little_endian 1.0", f"element vertex {len(pts)}", "property double x", "property double y", "property double z"]
    if cols is not None:
        fields += [("red", "u1"), ("green", "u1"), ("blue", "u1")]
        header += ["property uchar red", "property uchar green", "property uchar blue"]
    rec = np.zeros(len(pts), dtype=fields)
    rec["x"], rec["y"], rec["z"] = pts[:, 0], pts[:, 1], pts[:, 2]
    if cols is not None:
        c = np.clip(np.asarray(cols, dtype=np.float64) * 255.0, 0, 255).astype(np.uint8)
        rec["red"], rec["green"], rec["blue"] = c[:, 0], c[:, 1], c[:, 2]
    with open(path, "wb") as fp:
        fp.write(("\n".join(header + ["end_header"]) + "\n").encode("ascii"))
        fp.write(rec.tobytes())
    return True


io = types.SimpleNamespace(read_point_cloud=_read_point_cloud, write_point_cloud=_write_point_cloud)
visualization = types.SimpleNamespace(draw_geometries=lambda *a, **k: None)
