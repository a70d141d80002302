"""ctypes binding of libf3d.so (the C ABI in include/f3d.h).  There is no fallback: if the library or a CUDA
device is missing, every operator raises."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np
import torch

PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("F3D_LIB", PKG / "libf3d.so"))   # F3D_LIB: kernel-variant experiments only

NSTATS = 8
STAT_NAMES = ("candidates", "exact", "diverged", "near_edge", "seen", "audit_bad")
DEPTH_U16_MM, DEPTH_F32_M = 0, 1
FRAMES_U32, FRAMES_U32_T16 = 2, 3   # packed depth | class << 16 texels (f3d_pack_frames): row-major / 16x16 tiles

_vp, _i32, _i64, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double

# name -> (restype, argtypes); must list every symbol include/f3d.h declares (tests check this)
SIGNATURES = {
    "f3d_last_error": (C.c_char_p, []),
    "f3d_version": (C.c_int, []),
    "f3d_fuse_time_next_call": (C.c_int, [_vp, _vp]),
    "f3d_packed_frame_texels": (_i64, [_i32, _i32, _i32]),
    "f3d_pack_frames": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "f3d_frame_table_bytes": (_i64, [_i32]),
    "f3d_frames_setup": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _i32, _f64, _vp, _vp]),
    "f3d_frames_export": (C.c_int, [_vp, _i32, _vp, _vp, _vp, _vp]),
    "f3d_fuse_workspace_bytes": (_i64, [_i64]),
    "f3d_fuse_project_vote": (C.c_int, [_vp, _i64, _vp, _i32, _i32, _vp, _i32, _vp, _i32, _i32, _vp, _f64, _f64, _f64,
                                        _vp, _i32, _i32, _vp, _i64, _vp, _i32, _vp]),
    "f3d_fuse_project_vote_u16": (C.c_int, [_vp, _i64, _vp, _i32, _i32, _vp, _i32, _vp, _i32, _i32, _vp, _f64, _f64, _f64,
                                            _vp, _i32, _i32, _vp, _i64, _vp, _i32, _vp]),
    "f3d_resolve_labels_u16": (C.c_int, [_vp, _i64, _i32, _f64, _vp, _i32, _i32, _vp, _vp]),
    "f3d_exchange_constants": (C.c_int, [_vp]),
    "f3d_fuse_project_vote_exchange": (C.c_int, [_vp, _i64, _vp, _i32, _i32, _vp, _i32, _vp, _i32, _i32, _vp, _f64, _f64, _f64,
                                                 _i32, _i32, _i64, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _i64, _vp, _i32, _vp]),
    "f3d_exchange_publish": (C.c_int, [_vp, _vp, _i32, _i32, _i64, _vp]),
    "f3d_exchange_merge": (C.c_int, [_vp, _vp, _i32, _i64, _i64, _i64, _i32, _f64, _vp, _i32, _i32, _vp, _vp, _vp, _i64, _vp]),
    "f3d_exchange_queue_apply": (C.c_int, [_vp, _vp, _i32, _i64, _vp, _i64, _i32, _f64, _vp, _i32, _i32, _vp, _vp, _i64, _vp]),
    "f3d_fuse_project_vote_resolve": (C.c_int, [_vp, _i64, _vp, _i32, _i32, _vp, _i32, _vp, _i32, _i32, _vp, _f64, _f64,
                                                _f64, _vp, _i32, _f64, _vp, _i32, _i32, _vp, _vp, _i64, _vp, _i32, _vp]),
    "f3d_fuse_uv2pt": (C.c_int, [_vp, _i64, _vp, _i32, _i32, _vp, _i32, _i32, _i32, _vp, _f64, _f64, _f64, _vp, _vp,
                                 _i64, _vp, _i32, _vp]),
    "f3d_zbuffer_splat": (C.c_int, [_vp, _i64, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _i32, _vp, _i64, _vp, _i32,
                                    _vp]),
    "f3d_vote_uv2pt": (C.c_int, [_vp, _vp, _i32, _i64, _i32, _vp, _i64, _i32, _vp]),
    "f3d_vote_finalize": (C.c_int, [_vp, _i64, _vp]),
    "f3d_resize_nearest_u8": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _i32, _i32, _vp]),
    "f3d_resolve_labels": (C.c_int, [_vp, _i64, _i32, _f64, _vp, _i32, _i32, _vp, _vp]),
    "f3d_project_pixels": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "f3d_quat_rotate": (C.c_int, [_vp, _i64, _vp, _vp, _vp]),
    "f3d_frustum_mask": (C.c_int, [_vp, _i64, _vp, _vp, _i32, _vp, _vp]),
    "f3d_box_pairs_aabb": (C.c_int, [_vp, _vp, _vp, _i32, _vp, _i64, _vp, _vp]),
    "f3d_box_pairs_sweep": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _vp, _i64, _vp, _vp]),
    "f3d_union_find": (C.c_int, [_i32, _vp, _i64, _vp, _vp]),
    "f3d_obb_fit_workspace_bytes": (_i64, [_i32]),
    "f3d_obb_fit": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp]),
    "f3d_radius_grid_keys": (C.c_int, [_vp, _i64, _vp, _vp, _f64, _vp, _vp]),
    "f3d_radius_adjacency": (C.c_int, [_vp, _i64, _vp, _vp, _f64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "f3d_obb_contains": (C.c_int, [_vp, _i64, _vp, _i32, _vp, _vp]),
}

_lib = None


class F3dError(RuntimeError):
    pass


def load(build_if_missing: bool = True):
    """Load libf3d.so (building it in-tree with nvcc when absent).  Raises if neither is possible."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        if not build_if_missing:
            raise F3dError(f"{LIB_PATH} is missing: run __graft_entry__.build() (there is no CPU fallback)")
        from . import build as _build
        _build.build()
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().f3d_last_error()
        raise F3dError(f"libf3d {what} failed ({rc}): {msg.decode() if msg else ''}")


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise F3dError("no CUDA device: the B200 label-fusion path has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def ptr(t) -> int:
    """Device (torch tensor) or host (numpy array) address as an int for a void* argument; None -> NULL."""
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        assert t.is_contiguous(), "tensor must be contiguous"
        return t.data_ptr()
    if isinstance(t, np.ndarray):
        assert t.flags["C_CONTIGUOUS"], "array must be contiguous"
        return t.ctypes.data
    raise TypeError(type(t))


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def host_f64(a, n=None) -> np.ndarray:
    out = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))
    if n is not None and out.size != n:
        raise ValueError(f"expected {n} values, got {out.size}")
    return out
