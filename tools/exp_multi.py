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
shard = "contiguous" if "contiguous" in sys.argv else "interleaved"
fl, pts, K, wxyz, t, depth, masks = bench.build_scene(scenes, engine, fused, spec, 0, 0, torch, frame_ids=parallel.frame_shard_ids(spec.nframes, rank, world, shard))
N, C1 = fl.N, 134
what = (set(sys.argv[1:]) - {"contiguous", "interleaved"}) or {"records", "dense"}
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
if "records" in what:
    sx = parallel.VoteExchange(N, C1, torch.device("cuda", lr))
    def sx_fuse(**xa): engine.fuse_project_vote_exchange(fl.points4, fl.table, depth, masks, C1, radius=0.05, zmin=0.1, zmax=spec.zmax, **xa)
    def sx_fuse_only():
        sx.cursors.zero_(); sx_fuse(**sx.fuse_args())
    res["records: fuse + remote records"] = timeit(sx_fuse_only)
    engine.exchange_publish(sx.cursors, sx.peer_count_ptrs, rank, sx.sub_cap); torch.cuda.synchronize(); dist.barrier()
    res["records: merge (shard + labels)"] = timeit(lambda: engine.exchange_merge(sx.rx_slots, sx.rx_dir, world, sx.sub_rows, sx.per, sx.rows, C1, 133, 0.5, None, votes=sx.shard, labels=sx.lab))
    res["records: queue apply"] = timeit(lambda: engine.exchange_queue_apply(sx.rx_queue, sx.rx_count, world, sx.sub_cap, sx.shard, sx.rows, 133, sx.lab, 0.5, None))
    res["records: 2 barriers"] = timeit(lambda: (sx.hdl.barrier(channel=0), sx.hdl.barrier(channel=1)))
    res["records: all-gather labels"] = timeit(lambda: dist.all_gather_into_tensor(sx.full, sx.lab))
    res["records: whole step"] = timeit(lambda: sx.run(sx_fuse, 133, 0.5, None), reps=10)
    labels["records"] = sx.run(sx_fuse, 133, 0.5, None).clone()
    cur = sx.cursors.view(2 if False else 1, -1)[0]
    if rank == 0:
        rows = sx.cursors[:world * sx.nreg].view(world, sx.nreg).sum(dim=1).tolist(); q = sx.cursors[world * sx.nreg:].view(world, sx.nsub).sum(dim=1).tolist()
        print("records: rows per destination", rows, "max sub-region fill", int(sx.cursors[:world * sx.nreg].max()), "of", sx.sub_rows,
              "; queue entries per destination", q, "max sub-queue fill", int(sx.cursors[world * sx.nreg:].max()), "of", sx.sub_cap, flush=True)
    sx.check_overflow()
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
