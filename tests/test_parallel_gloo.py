"""CPU, world_size 2, gloo: the frame-sharded multi-GPU host logic (frame partition, chunked vote reduce-scatter,
shard-wise resolve, label all-gather) gives exactly the single-process result.  The per-rank partial votes come
from the oracle here (no GPU in this container); on the GPU box the same `fuse_sharded` driver wraps the CUDA
kernels (bench.py --gpus N)."""
import importlib
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
PKG_NAME = "3d-point-cloud-segmentation-using-2d-img-segmentation_b200"


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir, nchunks):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        parallel = importlib.import_module(PKG_NAME + ".parallel")
        scenes = importlib.import_module(PKG_NAME + ".scenes")
        from conftest import small_scene
        from oracle import f3d_oracle as orc
        s = small_scene(scenes, orc, npoints=6001, nframes=5, width=96, height=72, seed=61, block=8)
        a, b = parallel.frame_shard(len(s["t"]), rank, world)
        part = orc.fuse_project_vote(s["points"], s["K"], s["W"], s["H"], s["wxyz"], s["t"], s["depths"], s["masks"], 134, 0,
                                     0.05, 0.1, 4.0, 4.0, frames=range(a, b))
        part_t = torch.as_tensor(part.astype(np.int32))

        def fuse_chunk(lo, hi):
            return part_t[lo:hi].contiguous()

        def resolve(v):
            return torch.as_tensor(orc.segment(v.numpy(), 133, 0.5, None))

        labels = parallel.fuse_sharded(fuse_chunk, resolve, len(s["points"]), nchunks, "cpu")
        np.save(Path(out_dir) / f"labels_{rank}.npy", labels.numpy())
        if rank == 0:
            full = orc.fuse_project_vote(s["points"], s["K"], s["W"], s["H"], s["wxyz"], s["t"], s["depths"], s["masks"], 134,
                                         0, 0.05, 0.1, 4.0, 4.0)
            np.save(Path(out_dir) / "ref.npy", orc.segment(full, 133, 0.5, None))
        # the persistent pipeline object, run twice (buffers are reused between steps)
        pipe = parallel.ShardedPipeline(len(s["points"]), 134, nchunks, "cpu")
        for _ in range(2):
            lab2 = pipe.run(lambda a, b, out: out[:b - a].copy_(part_t[a:b]),
                            lambda v, out: out.copy_(torch.as_tensor(orc.segment(v.numpy(), 133, 0.5, None))))
        np.save(Path(out_dir) / f"labels2_{rank}.npy", lab2.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nchunks", [1, 3])
def test_frame_sharded_pipeline_world2(tmp_path, nchunks):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), nchunks), nprocs=world, join=True)
    ref = np.load(tmp_path / "ref.npy")
    assert (ref != 133).sum() > 100
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"labels_{r}.npy"), ref)
        assert np.array_equal(np.load(tmp_path / f"labels2_{r}.npy"), ref)


def test_shard_arithmetic():
    parallel = importlib.import_module(PKG_NAME + ".parallel")
    for n in (0, 1, 7, 500, 5000):
        for w in (1, 2, 3, 8):
            spans = [parallel.frame_shard(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def test_frame_and_point_sharding_helpers():
    """Host-side sharding rules of the multi-GPU path (no GPU needed)."""
    import importlib
    parallel = importlib.import_module(PKG_NAME + ".parallel")
    for nframes, world in [(500, 1), (4000, 8), (9, 8), (3, 8), (1001, 4)]:
        for mode in ("interleaved", "contiguous"):
            ids = [parallel.frame_shard_ids(nframes, r, world, mode) for r in range(world)]
            flat = sorted(i for part in ids for i in part)
            assert flat == list(range(nframes))                                  # a partition of the frames
            assert max(len(p) for p in ids) - min(len(p) for p in ids) <= 1      # balanced
        assert parallel.frame_shard_ids(nframes, 0, world, "contiguous") == list(range(*parallel.frame_shard(nframes, 0, world)))
    for n, world in [(10_000_000, 8), (20011, 3), (255, 2), (1, 8)]:
        per = parallel.shard_points(n, world)
        assert per % 256 == 0 and per * world >= n and (per - 256) * world < max(n, 256 * world)
