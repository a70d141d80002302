"""Golden vectors for the per-frame cache loader: the UNMODIFIED reference `FrameData` (`/root/reference/Fusion3DSeg/fusion.py:17-64`,
imported with the oracle/refshim stand-ins) reads a tiny RTAB cache written here and its `__getitem__` outputs are stored
(valid masks for decimation 1 / 2 / 3, frame names).  Output: tests/golden/g8_framedata.npz (inputs + outputs).
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_framedata.py
"""
import os
import pickle
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
REF = Path(os.environ.get("F3D_REFERENCE", "/root/reference"))
sys.dont_write_bytecode = True


def write_cache(root, depth_mm, frame_numbers):
    """Minimal RTAB cache: tofsegment list + per-frame pickles with the keys FrameData reads (`fusion.py:31-38`)."""
    mr = Path(root) / "PointcloudMergeResults" / "Segments_x"
    mr.mkdir(parents=True)
    F, h, w = depth_mm.shape
    rng = np.random.default_rng(1)
    tof = []
    for i in range(F):
        z = depth_mm[i].reshape(-1) / 1000
        org = np.stack([rng.normal(size=h * w), rng.normal(size=h * w), z], axis=1)
        data = {"frameNumber": int(frame_numbers[i]), "orgPoints": org, "modPoints": org + 1.0,
                "modSurfaceNormals": np.tile([0.0, 0.0, 1.0], (h * w, 1)), "orgColorPoints": np.zeros((h * w, 3), np.uint8)}
        rel = os.path.join("PointcloudMergeResults", "Segments_x", f"tofcameradata_segments_x_{i}.pkl")
        with open(Path(root) / rel, "wb") as fp:
            pickle.dump(data, fp)
        tof.append({"frameNumber": int(frame_numbers[i]), "fileName": rel + " "})
    tofp = Path(root) / "PointcloudMergeResults" / "tofsegment_x.pkl"
    with open(tofp, "wb") as fp:
        pickle.dump(tof, fp)
    return str(tofp)


def main():
    sys.path.insert(0, str(ROOT / "oracle" / "refshim"))
    sys.path.insert(0, str(REF))
    from Fusion3DSeg.fusion import FrameData as RefFrameData      # the reference, only needed to GENERATE the vectors
    rng = np.random.default_rng(7)
    h, w, F = 12, 16, 3
    depth = rng.integers(0, 6000, (F, h, w)).astype(np.uint16)
    depth[:, :2, :] = 0
    depth[0, 5, 5] = 100        # z == 0.1: excluded (strict >)
    depth[0, 5, 6] = 4000       # z == 4.0: included (<=)
    frames = np.array([3, 10, 11])
    out = {"depth_mm": depth, "frame_numbers": frames}
    with tempfile.TemporaryDirectory() as td:
        tof = write_cache(td, depth, frames)
        for dec in (1, 2, 3):
            fd = RefFrameData(tof, (0.1, 4), dec, (h, w))
            names, valids = [], []
            for i in range(len(fd)):
                name, pts, nrm, clr, valid = fd[i]
                names.append(name)
                valids.append(valid)
            out[f"valid_dec{dec}"] = np.stack(valids)
            out["names"] = np.array(names)
    np.savez_compressed(HERE / "g8_framedata.npz", **out)
    print("wrote g8_framedata.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
