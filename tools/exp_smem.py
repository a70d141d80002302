"""Experiment (1 GPU): dense-mode fused kernel time against extra dynamic shared memory (shrinks the L1 carve-out)."""
import importlib, os, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench
PKG = bench.PKG_NAME
engine = importlib.import_module(PKG + ".engine"); scenes = importlib.import_module(PKG + ".scenes"); fused = importlib.import_module(PKG + ".fused")
spec = scenes.CONFIGS["C2"]
fl, pts, K, wxyz, t, depth, masks = bench.build_scene(scenes, engine, fused, spec, 0, spec.nframes, torch)
votes = torch.empty((fl.N, 134), dtype=torch.int32, device="cuda")
def timed(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
for extra in (0, 8192, 16384, 24576, 28672, 32768):
    os.environ["F3D_EXTRA_SMEM"] = str(extra)
    print(f"extra smem {extra:6d} B: votes only {timed(lambda: engine.fuse_project_vote(fl.points4, fl.table, depth, masks, 134, 0.05, 0.1, spec.zmax, votes=votes)):7.3f} ms", flush=True)
