"""Multi-GPU timing experiment (torchrun): where does a frame-sharded step spend its time?  Components of the
slot-record exchange timed separately, the whole step for each exchange kind, and a cross-check of their labels."""
import importlib, os, sys
from pathlib import Path
import torch, torch.distributed as dist
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
what = set(sys.argv[1:]) or {"slots", "sparse", "dense"}
def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    tt = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    return float(tt.item())
res, labels = {}, {}
votes = torch.empty((N, C1), dtype=torch.int32, device="cuda")
res["single-GPU style fuse (dense votes)"] = timeit(lambda: engine.fuse_project_vote(fl.points4, fl.table, depth, masks, C1, 0.05, 0.1, spec.zmax, votes=votes))
del votes
if "slots" in what:
    sx = parallel.SlotExchange(N, C1, torch.device("cuda", lr))
    def sx_fuse(**xa): engine.fuse_project_vote_sparse(fl.points4, fl.table, depth, masks, C1, radius=0.05, zmin=0.1, zmax=spec.zmax, **xa)
    def sx_fuse_only():
        sx.cursors.zero_(); sx_fuse(**sx.fuse_args())
    res["slots: fuse + remote records"] = timeit(sx_fuse_only)
    engine.sparse_publish(sx.cursors, sx.peer_count_ptrs, rank, sx.cap); torch.cuda.synchronize(); dist.barrier()
    res["slots: merge (shard + labels)"] = timeit(lambda: engine.slots_merge(sx.rx_slots, sx.rx_dir, world, sx.rows_cap, sx.per, sx.rows, C1, 133, 0.5, None, votes=sx.shard, labels=sx.lab))
    res["slots: queue accumulate"] = timeit(lambda: engine.sparse_accumulate(sx.rx_queue, sx.rx_count, world, sx.cap, sx.shard, nrows=sx.rows))
    res["slots: queue relabel"] = timeit(lambda: engine.sparse_relabel(sx.rx_queue, sx.rx_count, world, sx.cap, sx.shard, sx.rows, 133, sx.lab, 0.5, None))
    res["slots: 2 barriers"] = timeit(lambda: (sx.hdl.barrier(channel=0), sx.hdl.barrier(channel=1)))
    res["slots: all-gather labels"] = timeit(lambda: dist.all_gather_into_tensor(sx.full, sx.lab))
    res["slots: whole step"] = timeit(lambda: sx.run(sx_fuse, 133, 0.5, None), reps=10)
    labels["slots"] = sx.run(sx_fuse, 133, 0.5, None).clone()
    if rank == 0: print("slots: cursors [queue x G, record rows x G]", sx.cursors.tolist(), "per", sx.per, "queue cap", sx.cap, "rows cap", sx.rows_cap, flush=True)
    sx.check_overflow()
if "sparse" in what:
    sp = parallel.SparseExchange(N, C1, torch.device("cuda", lr))
    def sparse_step():
        return sp.run(lambda q, cap, per, cur, ovf: engine.fuse_project_vote_sparse(fl.points4, fl.table, depth, masks, C1, q, cap, per, cur, ovf, 0.05, 0.1, spec.zmax),
                      lambda v, out: engine.resolve_labels(v, 133, 0.5, None, out=out))
    res["sparse: whole step"] = timeit(sparse_step)
    labels["sparse"] = sparse_step().clone()
    sp.check_overflow()
if "dense" in what:
    pipe = parallel.ShardedPipeline(N, C1, 4, torch.device("cuda", lr))
    def dense_step():
        return pipe.run(lambda a, b, out: engine.fuse_project_vote(fl.points4[a:b], fl.table, depth, masks, C1, 0.05, 0.1, spec.zmax, votes=out[:b - a]),
                        lambda v, out: engine.resolve_labels(v, 133, 0.5, None, out=out))
    res["dense packed reduce-scatter: whole step"] = timeit(dense_step)
    labels["dense"] = dense_step().clone()
torch.cuda.synchronize()
keys = sorted(labels)
for k in keys[1:]:
    assert torch.equal(labels[k], labels[keys[0]]), f"{k} exchange disagrees with {keys[0]}"
if rank == 0:
    for k, v in res.items(): print(f"{k:44s} {v:8.3f} ms", flush=True)
    print("labels agree across", keys, "checksum", int(labels[keys[0]].sum()), flush=True)
dist.destroy_process_group()
