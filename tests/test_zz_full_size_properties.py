"""BASELINE.json's metric configuration at FULL size (C2: 10 M points x 500 frames 1920x1440) through size-independent
properties -- the oracle cannot finish this size in seconds, so bit-exactness is carried here by checksums, linearity over
frames, idempotence and the equality of the two label paths (small-size tests compare with the oracle cell by cell).
Runs last (file name) and takes about half a minute on a B200."""
import importlib

import pytest

from conftest import PKG_NAME

pytestmark = pytest.mark.gpu


def test_c2_full_size_properties(engine, scenes):
    import torch

    import bench
    fused = importlib.import_module(PKG_NAME + ".fused")
    spec = scenes.CONFIGS["C2"]
    fl, pts, K, wxyz, t, depth, masks = bench.build_scene(scenes, engine, fused, spec, 0, spec.nframes, torch)
    N, F, C1 = fl.N, fl.table.F, 134
    assert (N, F) == (10_000_000, 500) and tuple(depth.shape) == (500, 1440, 1920)
    st = engine.new_stats()
    votes, labels = engine.fuse_project_vote_resolve(fl.points4, fl.table, depth, masks, C1, 133, 0.05, 0.1, spec.zmax, 0.5,
                                                     None, stats=st)
    torch.cuda.synchronize()
    s = engine.stats_dict(st)
    # checksum of checksums: every vote that was cast is in the tensor, nothing else is
    assert int(votes.sum(dtype=torch.int64)) == s["seen"] > 50_000_000
    assert int(votes.min()) >= 0 and int(votes.max()) <= F and s["audit_bad"] == 0
    assert s["exact"] < 0.02 * s["candidates"] and s["diverged"] <= s["exact"]
    # the fused epilogue's labels are exactly VotingSegmentation.segment of the written vote tensor (kernel 3)
    assert torch.equal(engine.resolve_labels(votes, 133, 0.5, None), labels)
    assert torch.equal(engine.resolve_labels(votes, 133, 0.3, [86, 114, 115]),
                       engine.fuse_project_vote_resolve(fl.points4, fl.table, depth, masks, C1, 133, 0.05, 0.1, spec.zmax, 0.3,
                                                        [86, 114, 115], want_votes=False)[1])
    # linearity over frames: the two halves of the frame set accumulate to the whole (votes commute, voting.py:98)
    half = F // 2
    acc = engine.fuse_project_vote(fl.points4, fl.table, depth[:half], masks[:half], C1, 0.05, 0.1, spec.zmax, frame_begin=0,
                                   frame_end=half)
    acc = engine.fuse_project_vote(fl.points4, fl.table, depth[half:], masks[half:], C1, 0.05, 0.1, spec.zmax, votes=acc,
                                   accumulate=True, frame_begin=half, frame_end=F)
    assert torch.equal(acc, votes)
    del acc
    # idempotence: the same launch again gives the same bytes (no dependence on atomics order or queue order)
    votes_b, labels_b = engine.fuse_project_vote_resolve(fl.points4, fl.table, depth, masks, C1, 133, 0.05, 0.1, spec.zmax, 0.5,
                                                         None)
    assert torch.equal(votes_b, votes) and torch.equal(labels_b, labels)
    del votes_b, labels_b
    # the ingest path's packed device layout (one uint32 texel per pixel, 16x16 tiles) gives the same bytes
    votes_p, labels_p = engine.fuse_project_vote_resolve(fl.points4, fl.table, fl.frames, None, C1, 133, 0.05, 0.1, spec.zmax, 0.5, None)
    assert torch.equal(votes_p, votes) and torch.equal(labels_p, labels)
