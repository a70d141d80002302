"""Multi-GPU timing experiment (torchrun): where does a frame-sharded step spend its time?"""
import importlib, os, sys, time
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
base = scenes.CONFIGS["C2"]
spec = scenes.scaled_spec("C2", nframes=base.nframes * world)
lo, hi = parallel.frame_shard(spec.nframes, rank, world)
fl, pts, K, wxyz, t, depth, masks = bench.build_scene(scenes, engine, fused, spec, lo, hi, torch)
N, C1 = fl.N, 134
def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    tt = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    return float(tt.item())
votes = torch.empty((N, C1), dtype=torch.int32, device="cuda")
def fuse_all(): engine.fuse_project_vote(fl.points4, fl.table, depth, masks, C1, 0.05, 0.1, spec.zmax, votes=votes)
def fuse_chunks(nch=8):
    for i in range(nch):
        a, b = i * N // nch, (i + 1) * N // nch
        engine.fuse_project_vote(fl.points4[a:b], fl.table, depth, masks, C1, 0.05, 0.1, spec.zmax, votes=votes[a:b])
out = torch.empty((N // world, C1), dtype=torch.int32, device="cuda")
def rs_all(): dist.reduce_scatter_tensor(out, votes, op=dist.ReduceOp.SUM)
v16 = torch.empty((N, C1 // 2), dtype=torch.int32, device="cuda"); o16 = torch.empty((N // world, C1 // 2), dtype=torch.int32, device="cuda")
def rs_half(): dist.reduce_scatter_tensor(o16, v16, op=dist.ReduceOp.SUM)
def resolve_shard(): engine.resolve_labels(out, 133, 0.5, None)
pipes = {}
def full(nch=8, packed=True):
    key = (nch, packed)
    if key not in pipes: pipes[key] = parallel.ShardedPipeline(N, C1, nch, torch.device("cuda", lr), packed=packed)
    return pipes[key].run(lambda a, b, out: engine.fuse_project_vote(fl.points4[a:b], fl.table, depth, masks, C1, 0.05, 0.1, spec.zmax, votes=out[:b - a]),
                          lambda v, out: engine.resolve_labels(v, 133, 0.5, None, out=out))
res = {"fuse 1 launch": timeit(fuse_all), "fuse 8 chunks": timeit(fuse_chunks), "reduce_scatter int32 5.36GB": timeit(rs_all),
       "reduce_scatter 2.68GB": timeit(rs_half), "resolve shard": timeit(resolve_shard), "pipeline 8 chunks": timeit(full),
       "pipeline 4 chunks": timeit(lambda: full(4)), "pipeline 16 chunks": timeit(lambda: full(16)),
       "pipeline 8 chunks int32": timeit(lambda: full(8, False))}
sp = parallel.SparseExchange(N, C1, torch.device("cuda", lr))
def sparse_step():
    return sp.run(lambda q, cap, per, cur, ovf: engine.fuse_project_vote_sparse(fl.points4, fl.table, depth, masks, C1, q, cap, per, cur, ovf, 0.05, 0.1, spec.zmax),
                  lambda v, out: engine.resolve_labels(v, 133, 0.5, None, out=out))
def sp_fuse_only():
    sp.cursors.zero_()
    engine.fuse_project_vote_sparse(fl.points4, fl.table, depth, masks, C1, sp.peer_queue_ptrs, sp.cap, sp.per, sp.cursors, sp.overflow, 0.05, 0.1, spec.zmax)
def sp_accum_only():
    sp.shard.zero_(); engine.sparse_accumulate(sp.rx, sp.rx_count, world, sp.cap, sp.shard)
def sp_barriers():
    sp.hdl.barrier(channel=0); sp.hdl.barrier(channel=1)
res["sparse: fuse+emit only"] = timeit(sp_fuse_only)
engine.sparse_publish(sp.cursors, sp.peer_count_ptrs, rank, sp.cap); torch.cuda.synchronize(); dist.barrier()
res["sparse: memset+accumulate only"] = timeit(sp_accum_only)
res["sparse: 2 barriers"] = timeit(sp_barriers)
res["sparse: resolve shard"] = timeit(lambda: engine.resolve_labels(sp.shard, 133, 0.5, None, out=sp.lab))
res["sparse exchange"] = timeit(sparse_step)
if rank == 0: print(f"{'sparse exchange':32s} {res['sparse exchange']:8.3f} ms  cursors={sp.cursors.tolist()}", flush=True)
sp.check_overflow()
ref = full(8, False).clone(); torch.cuda.synchronize()
assert torch.equal(sparse_step(), ref), "sparse exchange disagrees with the dense pipeline"
assert torch.equal(full(8), ref) and torch.equal(full(16), ref), "packed / chunked pipelines disagree"
single = engine.resolve_labels(votes, 133, 0.5, None) if world == 1 else None
if rank == 0:
    for k, v in res.items(): print(f"{k:32s} {v:8.3f} ms", flush=True)
dist.destroy_process_group()
