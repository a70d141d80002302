"""Experiment (GPU box only): raw statistics slots of one fused launch on a bench workload, with an optional variant library."""
import importlib, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench
PKG = bench.PKG_NAME
engine = importlib.import_module(PKG + ".engine"); scenes = importlib.import_module(PKG + ".scenes"); fused = importlib.import_module(PKG + ".fused")
wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
spec = scenes.CONFIGS[wl]
fl, pts, K, wxyz, t, depth, masks = bench.build_scene(scenes, engine, fused, spec, 0, spec.nframes, torch)
st = engine.new_stats()
votes = engine.fuse_project_vote(fl.points4, fl.table, depth, masks, 134, 0.05, 0.1, spec.zmax, stats=st)
torch.cuda.synchronize()
print("stats slots:", st.cpu().tolist())
vp = (votes > 0).sum(dim=1)
print("votes per point: mean %.2f max %d; distinct classes per point: mean %.2f max %d; >8 distinct: %.3f" % (
    votes.sum(dim=1).float().mean().item(), int(votes.sum(dim=1).max()), vp.float().mean().item(), int(vp.max()), (vp > 8).float().mean().item()))
print("max single cell", int(votes.max()))
