"""One library, one frame format, a few launches on the resident scene (run under ncu).  usage: exp_one.py LIB FMT [reps]  (env WL=C2|C1|C4)"""
import ctypes, importlib, os, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench
PKG = bench.PKG_NAME
engine = importlib.import_module(PKG + ".engine"); scenes = importlib.import_module(PKG + ".scenes"); fused = importlib.import_module(PKG + ".fused")
_lib = importlib.import_module(PKG + "._lib")
path, fmt = sys.argv[1], int(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
spec = scenes.CONFIGS[os.environ.get("WL", "C2")]
fl, pts, K, wxyz, t, depth, masks = bench.build_scene(scenes, engine, fused, spec, 0, spec.nframes, torch)
N, C1 = fl.N, 134
votes = torch.empty((N, C1), dtype=torch.int32, device="cuda"); labels = torch.empty(N, dtype=torch.int64, device="cuda")
ws = engine.workspace(N, fl.points4.device)
pk = engine.pack_frames(depth, masks, fmt) if fmt >= 2 else None
lib = ctypes.CDLL(path if path != "default" else str(_lib.LIB_PATH))
fn = lib.f3d_fuse_project_vote_resolve; fn.restype, fn.argtypes = _lib.SIGNATURES["f3d_fuse_project_vote_resolve"]
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
for i in range(reps + 1):
    if i == 1: e0.record()
    rc = fn(fl.points4.data_ptr(), N, fl.table.table.data_ptr(), 0, fl.table.F, depth.data_ptr() if fmt < 2 else pk.texels.data_ptr(), fmt,
            masks.data_ptr() if fmt < 2 else None, fl.table.H, fl.table.W, fl.table.K.ctypes.data, 0.05, 0.1, spec.zmax, votes.data_ptr(), C1, 0.5, None, 0, 133,
            labels.data_ptr(), ws.data_ptr(), ws.numel(), None, 0, torch.cuda.current_stream().cuda_stream)
    assert rc == 0
e1.record(); torch.cuda.synchronize()
print(path, "fmt", fmt, "call ms", e0.elapsed_time(e1) / reps, "votes", int(votes.sum()), "labels", int(labels.sum()))
