"""Scan-directory ingest for the fused path: the readers of the reference's
`RTAB_utils/ios_rtab.py` (`RTAB2Cache`) restated for the data the label fusion consumes, and a streaming driver that
decodes depth / mask PNGs on host threads into pinned staging buffers while the previous chunk of frames is copied to the
device and fused.

Reference data contract (all `file:line` into `RTAB_utils/ios_rtab.py` unless noted):
  * `calibration.yaml` -- two header lines skipped, `camera_matrix.data` reshaped to 3x3 (`:13-28`);
  * pose text -- `np.genfromtxt(delimiter=" ")`, rows `[startf:stopf]`, columns 0 = timestamp, 1:4 = xyz,
    4:8 = quaternion (x, y, z, w), 8 = image id (`:49-68`); `parse_rts` re-orders to (w, x, y, z) (`fusion.py:71-72`);
  * intrinsics scaled by (Depth_W / RGB_W, Depth_H / RGB_H) on (fx, cx) / (fy, cy) (`:115-131,167`);
  * depth `<id>.png` uint16 millimetres, optionally times a mask that zeroes a 10-pixel border (`:97-113`);
  * masks `<id>.png` uint8 class ids at RGB resolution, nearest-resized to the depth size (`voting.py:66,93`).
Only what the fused path needs is built: no RGB decode, no per-frame point pickles (`getTofCameraData`), no normals.
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import cv2
import numpy as np
import torch
import yaml

from .. import engine
from ..fused import FusedLabeler


def read_intrinsic(clib_file) -> np.ndarray:
    """`RTAB2Cache.__getIntrinsic` (`:13-28`): 3x3 camera matrix of the RTAB calibration YAML."""
    with open(clib_file) as infile:
        for _ in range(2):
            infile.readline()
        data = yaml.safe_load(infile)
    return np.reshape(data['camera_matrix']['data'], (3, 3))


def read_odometry(pose_file, startf=None, stopf=None):
    """`RTAB2Cache.__readOdometry` (`:49-68`): (img_idx [n], odo_xyz [n,3], odo_xyzw [n,4] as (x,y,z,w), timestamps [n])."""
    pose = np.genfromtxt(pose_file, delimiter=" ")
    pose = np.atleast_2d(pose)[startf:stopf]
    return np.asarray(pose[:, 8]), np.asarray(pose[:, 1:4]), np.asarray(pose[:, 4:8]), np.asarray(pose[:, 0])


def resize_camera_matrix(intrinsic, scale_x, scale_y) -> np.ndarray:
    """`RTAB2Cache.__resize_camera_matrix` (`:115-131`)."""
    fx, fy, cx, cy = intrinsic[0, 0], intrinsic[1, 1], intrinsic[0, 2], intrinsic[1, 2]
    return np.array([[fx * scale_x, 0.0, cx * scale_x], [0., fy * scale_y, cy * scale_y], [0., 0., 1.0]])


def read_depth_png(path, padding=False) -> np.ndarray:
    """One frame of `RTAB2Cache.__readDepth` (`:97-113`): uint16 millimetres; `padding` zeroes a 10-pixel border (the
    reference multiplies by a float mask of ones with a zero border -- same values, kept as uint16 here)."""
    d = cv2.imread(str(path), cv2.IMREAD_UNCHANGED)
    if d is None:
        raise FileNotFoundError(path)
    if d.ndim != 2 or d.dtype != np.uint16:
        raise ValueError(f"{path}: depth must be a single-channel 16-bit PNG (millimetres)")
    if padding:
        d = d.copy()
        d[:10, :] = 0
        d[-10:, :] = 0
        d[:, :10] = 0
        d[:, -10:] = 0
    return d


def read_mask_png(path) -> np.ndarray:
    """`cv2.imread(mask, 0)` (`segUtils/voting.py:66`)."""
    m = cv2.imread(str(path), 0)
    if m is None:
        raise FileNotFoundError(path)
    return m


class RTAB2Cache:
    """The reference's constructor arguments (`:30-47`); reads lazily and only what the fused path consumes."""

    def __init__(self, data_path, rgb_dir, depth_dir, pose_file, startf=None, stopf=None, stepf=1, relative=False,
                 padding=False, calibration=None):
        if relative:
            raise NotImplementedError("relative poses (`__globalRT2Local`) are not part of the label-fusion path")
        self.data_path, self.rgb_path, self.depth_path, self.pose_file = data_path, rgb_dir, depth_dir, pose_file
        self.startf, self.stopf, self.stepf, self.padding = startf, stopf, stepf, padding
        self.intrinsic = read_intrinsic(calibration if calibration is not None else Path(data_path) / 'calibration.yaml')
        self.img_idx, self.odo_xyz, self.odo_wxyz, self.odo_timestamp = read_odometry(pose_file, startf, stopf)

    def depth_file(self, k) -> Path:
        return Path(self.depth_path) / f"{int(self.img_idx[k])}.png"

    def rts(self, rgb_res, depth_res=None) -> dict:
        """The `rtsCameraData` dict of `getTofCameraData` (`:262-268`), i.e. what `parse_rts` (`fusion.py:67-77`) reads."""
        if depth_res is None:
            depth_res = read_depth_png(self.depth_file(0)).shape
        Ks = resize_camera_matrix(self.intrinsic, depth_res[1] / rgb_res[1], depth_res[0] / rgb_res[0])
        return {"intrinsic": self.intrinsic, "intrinsicScaled": Ks, "odo_wxyz": self.odo_wxyz, "odo_xyz": self.odo_xyz,
                "RGB_res": tuple(rgb_res), "Depth_res": tuple(depth_res)}


def label_scan(points, cache: RTAB2Cache, mask_dir, rgb_res, point_range=(0.1, 4), radius=0.05, nclasses=133,
               threshold=0.5, filter_classes=None, chunk_frames=32, workers=8, return_labeler=False):
    """Fused labels of a fixed cloud straight from an RTAB export directory.

    Frames (depth PNG + mask PNG per pose) are decoded by `workers` host threads into two pinned staging buffers; while
    chunk k is copied to the device and fused (masks nearest-resized on the GPU with OpenCV's index rule), chunk k+1 is
    being decoded.  Returns (votes float64 [N, nclasses+1], classes int64 [N]) -- the pair `get3DSeg.segment` returns
    (`get3DSeg.py:110`); frames without a mask file are skipped, as the reference pairs files by stem (`voting.py:45-54`)."""
    mask_dir = Path(mask_dir)
    frames = [k for k in range(len(cache.img_idx)) if (mask_dir / f"{int(cache.img_idx[k])}.png").is_file()
              and cache.depth_file(k).is_file()]
    if not frames:
        raise FileNotFoundError("no frame has both a depth PNG and a mask PNG")
    h, w = read_depth_png(cache.depth_file(frames[0])).shape
    rts = cache.rts(rgb_res, (h, w))
    wxyz = np.ascontiguousarray(rts["odo_wxyz"][frames][:, [3, 0, 1, 2]])        # fusion.py:71-72
    trans = np.ascontiguousarray(rts["odo_xyz"][frames])
    fl = FusedLabeler(points, rts["intrinsicScaled"], w, h, wxyz, trans, point_range, radius, nclasses)
    mh, mw = read_mask_png(mask_dir / f"{int(cache.img_idx[frames[0]])}.png").shape
    cf = max(1, min(int(chunk_frames), len(frames)))
    dev = fl.points4.device
    hd = [torch.empty((cf, h, w), dtype=torch.uint16, pin_memory=True) for _ in range(2)]
    hm = [torch.empty((cf, mh, mw), dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    dd = [torch.empty((cf, h, w), dtype=torch.uint16, device=dev) for _ in range(2)]
    dm = [torch.empty((cf, mh, mw), dtype=torch.uint8, device=dev) for _ in range(2)]
    copied = [torch.cuda.Event() for _ in range(2)]
    fused_done = [torch.cuda.Event() for _ in range(2)]
    compute, copy = torch.cuda.current_stream(), torch.cuda.Stream(device=dev)

    def decode(slot, j, k):
        d = read_depth_png(cache.depth_file(k), cache.padding)
        m = read_mask_png(mask_dir / f"{int(cache.img_idx[k])}.png")
        if d.shape != (h, w) or m.shape != (mh, mw):
            raise ValueError(f"frame {int(cache.img_idx[k])}: image size differs from the first frame")
        hd[slot][j].copy_(torch.from_numpy(d))
        hm[slot][j].copy_(torch.from_numpy(m))

    chunks = [frames[a:a + cf] for a in range(0, len(frames), cf)]
    with ThreadPoolExecutor(max_workers=max(1, int(workers))) as pool:
        pending = [pool.submit(decode, 0, j, k) for j, k in enumerate(chunks[0])]
        for c, chunk in enumerate(chunks):
            s = c & 1
            for f in pending:
                f.result()
            if c + 1 < len(chunks):
                # the other slot's host buffers are free once its previous copy has completed
                if c >= 1:
                    copied[s ^ 1].synchronize()
                pending = [pool.submit(decode, s ^ 1, j, k) for j, k in enumerate(chunks[c + 1])]
            n = len(chunk)
            with torch.cuda.stream(copy):
                if c >= 2:
                    copy.wait_event(fused_done[s])          # device staging of this slot was read by chunk c-2's kernels
                dd[s][:n].copy_(hd[s][:n], non_blocking=True)
                dm[s][:n].copy_(hm[s][:n], non_blocking=True)
                copied[s].record(copy)
            compute.wait_event(copied[s])
            masks = dm[s][:n] if (mh, mw) == (h, w) else engine.resize_nearest(dm[s][:n], h, w)    # voting.py:93
            a = c * cf
            fl.vote(dd[s][:n], masks, frame_begin=a, frame_end=a + n)
            fused_done[s].record(compute)
    classes = fl.segment(threshold, filter_classes).cpu().numpy()
    votes = fl.votes_numpy()
    return (votes, classes, fl) if return_labeler else (votes, classes)
