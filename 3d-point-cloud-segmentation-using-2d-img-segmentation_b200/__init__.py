"""fusion3dseg-b200: B200-native (sm_100a) multi-view 2D->3D label fusion.

Drop-in for the label-fusion hot path of raviraj988/3D-POINT-CLOUD-SEGMENTATION-USING-2D-IMG-SEGMENTATION:
the sub-packages `Fusion3DSeg`, `RTAB_utils` and the module `get3DSeg` mirror the reference's import paths and
call signatures; the arithmetic runs in hand-written CUDA kernels behind the C ABI of include/f3d.h
(csrc/ -> libf3d.so).  No CPU fallback: operators raise `F3dError` without the library or a CUDA device.
"""
from . import _lib, engine, scenes  # noqa: F401
from ._lib import F3dError, load  # noqa: F401

__all__ = ["engine", "scenes", "load", "F3dError"]
