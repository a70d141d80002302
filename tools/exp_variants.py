"""Timing / equality experiment (GPU box only): every libf3d variant under build/variants/ (tools/build_variants.py) plus the
in-tree library on the same resident scene, for the two-array, packed row-major and packed tiled frame formats.
usage: python tools/exp_variants.py [C2|C1|C4] [reps]"""
import ctypes, glob, importlib, os, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench
PKG = bench.PKG_NAME
engine = importlib.import_module(PKG + ".engine"); scenes = importlib.import_module(PKG + ".scenes"); fused = importlib.import_module(PKG + ".fused")
_lib = importlib.import_module(PKG + "._lib")
wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
spec = scenes.CONFIGS[wl]
fl, pts, K, wxyz, t, depth, masks = bench.build_scene(scenes, engine, fused, spec, 0, spec.nframes, torch)
N, C1 = fl.N, 134
votes = torch.empty((N, C1), dtype=torch.int32, device="cuda"); labels = torch.empty(N, dtype=torch.int64, device="cuda")
ws = engine.workspace(N, fl.points4.device)
packed = {2: engine.pack_frames(depth, masks, 2), 3: engine.pack_frames(depth, masks, 3)}
torch.cuda.synchronize()
SIG = _lib.SIGNATURES["f3d_fuse_project_vote_resolve"]
def bind(path):
    lib = ctypes.CDLL(path)
    fn = lib.f3d_fuse_project_vote_resolve; fn.restype, fn.argtypes = SIG
    lib.f3d_last_error.restype = ctypes.c_char_p
    return lib
def run(lib, fmt, reps):
    dptr = depth.data_ptr() if fmt == 0 else packed[fmt].texels.data_ptr()
    mptr = masks.data_ptr() if fmt == 0 else None
    def call():
        rc = lib.f3d_fuse_project_vote_resolve(fl.points4.data_ptr(), N, fl.table.table.data_ptr(), 0, fl.table.F, dptr, fmt, mptr, fl.table.H, fl.table.W,
                                               fl.table.K.ctypes.data, 0.05, 0.1, spec.zmax, votes.data_ptr(), C1, 0.5, None, 0, 133, labels.data_ptr(),
                                               ws.data_ptr(), ws.numel(), None, 0, torch.cuda.current_stream().cuda_stream)
        assert rc == 0, lib.f3d_last_error()
    votes.zero_(); labels.zero_()
    for _ in range(2): call()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): call()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, (int(votes.sum()), int((votes.to(torch.int64) * torch.arange(C1, device="cuda")).sum()), int(labels.sum()))
ref = None
libs = sorted(glob.glob(str(ROOT / "build" / "variants" / "*.so")), key=lambda p: (not p.endswith("_r1.so"), p)) + [str(_lib.LIB_PATH)]
for path in libs:
    lib = bind(path)
    name = os.path.basename(path)
    for fmt in ((0,) if name.endswith("_r1.so") else (0, 2, 3)):
        try:
            ms, sig = run(lib, fmt, reps)
        except AssertionError as ex:
            print(f"{name:24s} fmt {fmt}: FAILED {ex}", flush=True); continue
        ref = ref or sig
        print(f"{name:24s} fmt {fmt}: call {ms:7.3f} ms  {'same' if sig == ref else 'DIFFERENT ' + str(sig) + ' vs ' + str(ref)}", flush=True)
