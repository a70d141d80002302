"""Where does a frame-sharded exchange step spend its time?  (torchrun, GPU box)  Phases of parallel.VoteExchange.run timed with CUDA
events on a C3-shaped scene.  usage: torchrun ... tools/exp_exchange_breakdown.py [points] [frames] [contiguous|interleaved]"""
import importlib, os, sys
from pathlib import Path
import numpy as np, torch, torch.distributed as dist
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench
PKG = bench.PKG_NAME
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
engine = importlib.import_module(PKG + ".engine"); scenes = importlib.import_module(PKG + ".scenes")
fused = importlib.import_module(PKG + ".fused"); parallel = importlib.import_module(PKG + ".parallel")
parallel.init_process_group(lr)
npts = int(sys.argv[1]) if len(sys.argv) > 1 else 25_000_000
nfr = int(sys.argv[2]) if len(sys.argv) > 2 else 1250
mode = sys.argv[3] if len(sys.argv) > 3 else "contiguous"
spec = scenes.scaled_spec("C3", npoints=npts, nframes=nfr)
p4 = torch.empty((npts, 4), dtype=torch.float32, device="cuda")
if rank == 0:
    p4[:, :3] = torch.as_tensor(scenes.make_cloud(spec)).cuda(); p4[:, 3] = 0
dist.broadcast(p4, 0)
ids = parallel.frame_shard_ids(nfr, rank, world, mode)
fl, K, wxyz, t = bench.make_labeler(fused, scenes, spec, ids, p4)
bench.build_frames(torch, engine, fl, spec, ids)
x = parallel.VoteExchange(npts, 134, torch.device("cuda", lr))
names = ["barrier0+zero", "supertile+fuse+fixup", "publish", "barrier1", "merge+label stores", "queue_apply", "barrier2+widen", "overflow"]
acc = np.zeros(len(names))
def step(record):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    ev[0].record()
    x.hdl.barrier(channel=0); x.cursors.zero_(); x.overflow.zero_(); ev[1].record()
    engine.fuse_project_vote_exchange(fl.points4, fl.table, fl.frames, None, 134, radius=0.05, zmin=fl.zmin, zmax=fl.zmax, **x.fuse_args()); ev[2].record()
    engine.exchange_publish(x.cursors, x.peer_count_ptrs, x.rank, x.sub_cap); ev[3].record()
    x.hdl.barrier(channel=1); ev[4].record()
    engine.exchange_merge(x.rx_slots, x.rx_dir, x.world, x.sub_rows, x.per, x.rows, x.c1, 133, 0.5, None, votes=x.shard, labels=x.lab, peer_labels16=x.peer_label_ptrs, first_point=x.rank * x.per); ev[5].record()
    engine.exchange_queue_apply(x.rx_queue, x.rx_count, x.world, x.sub_cap, x.shard, x.rows, 133, x.lab, 0.5, None, peer_labels16=x.peer_label_ptrs, first_point=x.rank * x.per); ev[6].record()
    x.hdl.barrier(channel=2); x.full.copy_(x.full16); ev[7].record()
    x.ovf_any.copy_(x.overflow); dist.all_reduce(x.ovf_any, op=dist.ReduceOp.MAX); ev[8].record()
    torch.cuda.synchronize()
    if record:
        for i in range(len(names)): acc[i] += ev[i].elapsed_time(ev[i + 1])
for _ in range(3): step(False)
dist.barrier()
n = 8
for _ in range(n): step(True)
tt = torch.tensor(acc / n, device="cuda"); mx = tt.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"world {world} points {npts} frames {nfr} ({mode}); per-rank rows {x.rows}; ms per phase (rank 0 | max over ranks)")
    for nm, a, b in zip(names, tt.tolist(), mx.tolist()): print(f"  {nm:28s} {a:8.3f} {b:8.3f}")
    print(f"  {'sum':28s} {sum(tt.tolist()):8.3f}")
    print("  queue entries rank0:", int(x.rx_count.view(torch.int32).sum()), " overflow", int(x.ovf_any.item()))
dist.destroy_process_group()
