"""How much of a frame-sharded rank's sweep is spent on tiles none of its frames can see?  (GPU box, one GPU)
The C3 cloud with the frames of one contiguous 1/8 shard, labels only (no dense vote write): kernel time and candidates
against the full 5000-frame launch.  usage: python tools/exp_dead_tiles.py [npoints] [nframes] [shards]"""
import importlib, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench
PKG = bench.PKG_NAME
engine = importlib.import_module(PKG + ".engine"); scenes = importlib.import_module(PKG + ".scenes"); fused = importlib.import_module(PKG + ".fused")
npts = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
nfr = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
G = int(sys.argv[3]) if len(sys.argv) > 3 else 8
spec = scenes.scaled_spec("C3", npoints=npts, nframes=nfr)
pts = torch.as_tensor(scenes.make_cloud(spec)).cuda()
p4 = torch.zeros((npts, 4), dtype=torch.float32, device="cuda"); p4[:, :3] = pts; del pts
tot = 0.0
for r in list(range(G)):
    a, b = r * nfr // G, (r + 1) * nfr // G
    ids = list(range(a, b))
    fl, K, wxyz, t = bench.make_labeler(fused, scenes, spec, ids, p4)
    bench.build_frames(torch, engine, fl, spec, ids)
    kt = engine.KernelTimer()
    for _ in range(2): fl.label(want_votes=False)
    fl.stats.zero_()
    ms = bench.timed(torch, lambda: fl.label(want_votes=False, timer=kt), 5, warmup=0)
    st = fl.stats_dict()
    nst = (npts + 4095) // 4096
    print(f"shard {r} frames [{a},{b}): call {ms:.3f} ms kernel {np.mean(kt.ms()):.3f} ms candidates/step {st['candidates'] / 5:.4g} seen/step {st['seen'] / 5:.4g}", flush=True)
    tot += float(np.mean(kt.ms()))
    del fl
    torch.cuda.empty_cache()
print(f"sum of shard kernels {tot:.2f} ms (the 5000-frame launch: configs.C3 labels_only_ms of the --gpus 1 line)")
