"""CPU: the C-ABI library builds, loads and exports exactly the symbols include/f3d.h declares; operators fail
loudly (no CPU fallback) when no CUDA device is present."""
import importlib
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from conftest import PKG_NAME, ROOT


def header_symbols():
    txt = (ROOT / "include" / "f3d.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(f3d_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.load()
    names = header_symbols()
    assert len(names) >= 17
    _lib = importlib.import_module(PKG_NAME + "._lib")
    assert sorted(_lib.SIGNATURES) == names
    for n in names:
        assert getattr(lib, n) is not None
    assert lib.f3d_version() >= 100
    assert lib.f3d_frame_table_bytes(10) == 10 * 656


def test_bad_arguments_are_reported(pkg):
    lib = pkg.load()
    rc = lib.f3d_resolve_labels(None, 10, 134, 0.5, None, 0, 133, None, None)
    assert rc == -1
    assert b"f3d_resolve_labels" in lib.f3d_last_error()
    K = np.eye(3)
    K[1, 0] = 0.5
    rc = lib.f3d_frames_setup(K.ctypes.data, 4, 4, 1, 1, 1, 4.0, 1, None)
    assert rc == -3 and b"upper triangular" in lib.f3d_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(pkg):
    cam = importlib.import_module(PKG_NAME + ".Fusion3DSeg.camera_utils")
    with pytest.raises(pkg.F3dError):
        cam.points2pixel(np.zeros((4, 3)), np.eye(3), [1, 0, 0, 0], [0, 0, 0])
    fused = importlib.import_module(PKG_NAME + ".fused")
    with pytest.raises(pkg.F3dError):
        fused.FusedLabeler(np.zeros((4, 3), np.float32), np.eye(3), 8, 8, [[1, 0, 0, 0]], [[0, 0, 0]])


def test_scene_generator_is_deterministic(scenes):
    spec = scenes.scaled_spec("C1", npoints=5000, nframes=5, width=64, height=48, seed=3)
    a, b = scenes.make_cloud(spec), scenes.make_cloud(spec)
    assert a.dtype == np.float32 and a.shape == (5000, 3) and np.array_equal(a, b)
    q, t = scenes.make_poses(spec)
    assert q.shape == (5, 4) and t.shape == (5, 3)
    assert np.array_equal(np.round(q, 6), q)                      # 6-decimal pose text
    assert np.abs((q ** 2).sum(1) - 1.0).max() > 1e-8            # therefore not unit quaternions
    K = scenes.scaled_intrinsics(1920, 1440)
    assert K[0, 0] == scenes.CALIB_FX * (1920 / 720) and K[1, 2] == scenes.CALIB_CY * (1440 / 960)
    m = scenes.block_masks((48, 64), 3, seed=1, block=8)
    assert m.shape == (3, 48, 64) and m.max() <= 133


def test_vote_epilogue_keeps_its_32_byte_stores(pkg):
    """ptxas 12.9 was seen to lower `st.global.v8.b32` to ONE 32-bit store when the vote flush was a non-inlined function
    (7 of 8 vote cells never written); the epilogue instance is inlined for that reason.  Every byte-histogram
    instantiation of the fused kernel must still contain the 32-byte store, and the in-tree library must hold sm_100a code
    for the production kernel."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    lib = ROOT / PKG_NAME / "libf3d.so"
    if not Path(cuobjdump).exists() or not lib.exists():
        pytest.skip("cuobjdump or libf3d.so not available")
    sass = subprocess.run([cuobjdump, "-sass", str(lib)], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    funcs = re.split(r"\n\s*Function : ", sass)
    seen = 0
    for f in funcs:
        m = re.match(r"_Z11fuse_kernelILi0ELi(\d)ELi1ELb([01])E", f)
        if m:
            seen += 1
            assert "STG.E.ENL2.256" in f, f"fuse_kernel<VOTE,{m.group(1)},HB1,{m.group(2)}> lost its 32-byte vote stores"
            assert "LDGSTS" in f                      # cp.async staging of the pose tiles (and texels for the packed formats)
    assert seen == 8
    assert "UBLKCP" in sass                           # TMA bulk copies stage the fix-up kernel's frame records
    assert "LTC64B" in sass                           # 64-byte L2 fills for the frame gathers
