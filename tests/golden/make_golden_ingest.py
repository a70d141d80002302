"""Golden vectors for the scan-directory ingest, produced by the UNMODIFIED reference readers.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_ingest.py

A tiny synthetic RTAB export (pose text in the `rtabmap-export` "RGBD-SLAM + ID" format, 16-bit depth PNGs) is written to
a temp dir and read back through the private methods of `RTAB_utils/ios_rtab.py:RTAB2Cache` (`__getIntrinsic` :13-28,
`__readOdometry` :49-68, `__readDepth` :97-113, `__resize_camera_matrix` :115-131).  The reference module needs
`open3d`, `pyquaternion` and `skimage` at import time: the stand-ins under `oracle/refshim` (never called here).
The constructor opens `RTAB_utils/calibration.yaml` relative to the working directory, so it runs with cwd = reference."""
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

import cv2  # noqa: E402


def main():
    rng = np.random.Generator(np.random.PCG64(606))
    n, H, W = 7, 48, 64
    ids = np.array([3, 4, 7, 8, 12, 13, 20])
    ts = 1.7e9 + np.arange(n) * 0.0333
    xyz = np.round(rng.normal(0, 2, (n, 3)), 6)
    q = rng.normal(size=(n, 4))
    q = np.round(q / np.linalg.norm(q, axis=1, keepdims=True), 6)              # (x, y, z, w), 6 decimals: not unit
    lines = ["%.6f %.6f %.6f %.6f %.6f %.6f %.6f %.6f %d" % (ts[i], *xyz[i], *q[i], ids[i]) for i in range(n)]
    pose_text = "#timestamp x y z qx qy qz qw id\n" + "\n".join(lines) + "\n"
    depths = rng.integers(0, 6000, (n, H, W)).astype(np.uint16)
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        (td / "depth").mkdir()
        (td / "rgb").mkdir()
        (td / "poses.txt").write_text(pose_text)
        for i in range(n):
            assert cv2.imwrite(str(td / "depth" / f"{ids[i]}.png"), depths[i])
        cwd = os.getcwd()
        os.chdir(REF)
        try:
            from RTAB_utils.ios_rtab import RTAB2Cache
            out = {}
            for tag, (a, b, pad) in {"all": (None, None, False), "slice_pad": (1, 6, True)}.items():
                c = RTAB2Cache(str(td), str(td / "rgb"), str(td / "depth"), str(td / "poses.txt"), a, b, 1, False, pad)
                img_idx, odo_xyz, odo_xyzw, stamp = c._RTAB2Cache__readOdometry()
                c.img_idx = img_idx
                d = c._RTAB2Cache__readDepth()
                out[f"{tag}_img_idx"], out[f"{tag}_odo_xyz"], out[f"{tag}_odo_xyzw"], out[f"{tag}_stamp"] = img_idx, odo_xyz, odo_xyzw, stamp
                out[f"{tag}_depths"] = np.stack([np.asarray(x, dtype=np.float64) for x in d])
                out[f"{tag}_range"] = np.array([-1 if a is None else a, -1 if b is None else b, int(pad)])
            out["intrinsic"] = c.intrinsic
            out["scaled_64x48_from_720x960"] = c._RTAB2Cache__resize_camera_matrix(64 / 720, 48 / 960)
            out["scaled_1920x1440_from_720x960"] = c._RTAB2Cache__resize_camera_matrix(1920 / 720, 1440 / 960)
        finally:
            os.chdir(cwd)
    out["pose_text"] = np.frombuffer(pose_text.encode(), dtype=np.uint8)
    out["calibration_text"] = np.frombuffer((REF / "RTAB_utils" / "calibration.yaml").read_bytes(), dtype=np.uint8)
    out["depths_u16"] = depths
    out["ids"] = ids
    np.savez_compressed(HERE / "g6_ingest.npz", **out)
    print({k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
