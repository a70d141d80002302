#!/usr/bin/env python
"""Benchmark of the fused label-fusion hot path (BASELINE.json metric: point-view projections / s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload C2|C1|C3|C4|small] [--configs C1,C4,C5,micro]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over the whole synthetic scene.

N = 1 (default workload C2, BASELINE configs[1]): ONE launch of kernel (1) -- fused project + z-test + frame gather + vote over
every frame with the label resolve (kernel 3's arithmetic) in its epilogue -- on frames resident in HBM in the ingest
path's packed device layout (uint32 texel = depth mm | class << 16, `f3d_pack_frames`).  The same line also reports the step
on the two-array inputs (uint16 depth + uint8 mask stacks) and with the re-pack inside the step, the end-to-end step from
pinned HOST buffers through the public API, the CPU port of the reference on the host cores (both figures), and nested
lines for the other BASELINE configs (C1 compared cell for cell with the CPU port, C4, C5) and the other kernels.

N > 1 (default workload C3, configs[2]: 100 M points x 5000 frames = a FIXED problem split over the ranks, strong scaling):
frames sharded over the ranks (contiguous = headline, interleaved and point-sharded also timed), the vote exchange fused into
the kernel (slot records written straight into the owner rank's memory over NVLink; launched only over the super-tiles a rank's
frames can see, `--compact`), owner-side merge + label resolve with the labels stored into every rank's copy from inside the
merge.  Rank 0 additionally runs the whole frame set on its single GPU and the CPU port on a sample, and the line says whether
the N-rank result equals both.

A point-view is one (point, frame) pair of the nominal N_points x N_frames product.  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import hashlib
import importlib
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
PKG_NAME = "3d-point-cloud-segmentation-using-2d-img-segmentation_b200"

METRIC = "point_view_projections_per_sec"
UNIT = "point-views/s"
NCLASSES = 133
C1 = NCLASSES + 1
RADIUS, THRESHOLD = 0.05, 0.5

WORKLOADS = {
    # name: (config key in scenes.CONFIGS, description)
    "C2": ("C2", "configs[1] iOS RTAB-style scan: 10M points x 500 frames 1920x1440 uint16-mm depth + uint8 masks"),
    "C1": ("C1", "configs[0] CPU-reference scene: 1M points x 50 frames 640x480"),
    "C3": ("C3", "configs[2] large building scan: 100M points x 5000 frames 1920x1440, frames sharded over the GPUs"),
    "C4": ("C4", "configs[3] dense 4K: 20M points x 1000 frames 3840x2160"),
    "small": ("C1", "debug scene: 200k points x 8 frames 320x240"),
}
T_START = time.time()


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                          str(gpu_index), "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                         text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.25)
        self.proc.terminate()
        rows = [r for ts, r in self.rows if t0 - 0.1 <= ts <= t1 + 0.3 and len(r) >= 9] or [r for _, r in self.rows if len(r) >= 9]
        if not rows:
            return None
        sm = sorted(float(r[1]) for r in rows)
        reasons = set()
        for r in rows:
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "reasons": sorted(reasons),
                "power_w_max": max(float(r[3]) for r in rows), "samples": len(rows)}


def algorithmic_bytes(N, F, H, W, depth_bytes, c1, rows=None):
    """SURVEY 8(d) designated figure for kernel (1): every input read once, every output written once (`rows`: vote rows this
    launch writes, default all N)."""
    return 16 * N + F * H * W * (1 + depth_bytes) + 64 * F + 4 * (N if rows is None else rows) * c1


def kernel_source_hash():
    h = hashlib.sha1()
    for f in ("fuse_kernel.cuh", "f3d_common.cuh"):
        h.update((ROOT / PKG_NAME / "csrc" / f).read_bytes())
    return h.hexdigest()[:12]


def profile_figures(workload):
    """ncu-derived figures of the fused kernel committed under profiles/ (DRAM traffic per launch, fp32 pipe utilisation).
    They are stamped with the hash of the kernel source they were captured on and dropped when the source has changed."""
    tj = ROOT / "profiles" / "fuse_kernel_traffic.json"
    if not tj.exists():
        return None
    try:
        d = json.loads(tj.read_text())
    except Exception:   # noqa: BLE001
        return None
    if d.get("kernel_source_hash") != kernel_source_hash() or d.get("workload") != workload:
        return None
    return d


def frame_masks(torch, frame_ids, H, W, seed, device="cuda", block=32):
    """uint8 [F,H,W] block-constant label images, a pure function of (seed, global frame id, block): every rank -- and the
    single-GPU reference run of the multi-GPU parity check -- produces identical masks for the same frame id.  Labels
    0..132, 5 % of the blocks = 133 (unclassified, `get2DSeg.py:118`)."""
    bh, bw = -(-H // block), -(-W // block)
    fid = torch.as_tensor(np.asarray(frame_ids, dtype=np.int64), device=device)[:, None, None]
    by = torch.arange(bh, device=device, dtype=torch.int64)[None, :, None]
    bx = torch.arange(bw, device=device, dtype=torch.int64)[None, None, :]
    h = (fid * 73856093 + by * 19349663 + bx * 83492791 + int(seed) * 2654435761) & 0x7fffffff
    h = ((h ^ (h >> 15)) * 0x2c1b3c6d) & 0x7fffffff
    h = ((h ^ (h >> 12)) * 0x297a2d39) & 0x7fffffff
    h = h ^ (h >> 15)
    lab = (h % NCLASSES).to(torch.uint8)
    lab[((h >> 8) % 20) == 0] = NCLASSES
    return lab.repeat_interleave(block, dim=1).repeat_interleave(block, dim=2)[:, :H, :W].contiguous()


def build_frames(torch, engine, fl, spec, frame_ids, keep_unpacked=False):
    """Depth = GPU z-buffer splat of the cloud (kernel 2) for the labeler's frames, masks = `frame_masks`; both are packed
    chunk by chunk into the labeler's resident packed stack.  Returns (depth, masks) stacks when `keep_unpacked`."""
    F, H, W = fl.nframes, spec.height, spec.width
    step = max(1, min(F, (1 << 30) // (H * W * 4)))
    zbuf = torch.empty((step, H * W), dtype=torch.int32, device="cuda")
    depth = torch.empty((F if keep_unpacked else step, H, W), dtype=torch.uint16, device="cuda")
    masks = torch.empty((F, H, W), dtype=torch.uint8, device="cuda") if keep_unpacked else None
    for a in range(0, F, step):
        b = min(a + step, F)
        d = depth[a:b] if keep_unpacked else depth[: b - a]
        engine.zbuffer_splat(fl.points4, fl.table, border=10, frame_begin=a, frame_end=b, zbuf=zbuf, out=d)
        for a2 in range(a, b, 32):
            b2 = min(a2 + 32, b)
            m = frame_masks(torch, frame_ids[a2:b2], H, W, spec.seed)
            if keep_unpacked:
                masks[a2:b2] = m
            fl.pack(d[a2 - a:b2 - a], m, frame_begin=a2)
    torch.cuda.synchronize()
    del zbuf
    return (depth, masks) if keep_unpacked else (None, None)


def make_labeler(fused, scenes, spec, frame_ids, pts):
    K = scenes.scaled_intrinsics(spec.width, spec.height)
    wxyz, t = scenes.make_poses(spec)
    ids = np.asarray(frame_ids, dtype=np.int64)
    fl = fused.FusedLabeler(pts, K, spec.width, spec.height, np.ascontiguousarray(wxyz[ids]), np.ascontiguousarray(t[ids]),
                            point_range=(0.1, spec.zmax), radius=RADIUS, nclasses=NCLASSES)
    return fl, K, wxyz, t


def build_scene(scenes, engine, fused, spec, frame_lo, frame_hi, torch, frame_ids=None):
    """Whole synthetic scene on the device (tests / tools): labeler with its packed frame stack plus the two-array stacks."""
    ids = list(range(frame_lo, frame_hi)) if frame_ids is None else list(frame_ids)
    pts = scenes.make_cloud(spec)
    fl, K, wxyz, t = make_labeler(fused, scenes, spec, ids, pts)
    depth, masks = build_frames(torch, engine, fl, spec, ids, keep_unpacked=True)
    return fl, pts, K, np.ascontiguousarray(wxyz[ids]), np.ascontiguousarray(t[ids]), depth, masks


def tensor_sum(t, rows=4_000_000):
    """Exact integer sum of a big int32 tensor without torch's full-size int64 temporary."""
    return sum(int(t[a:a + rows].sum()) for a in range(0, t.shape[0], rows))


def timed(torch, fn, steps, warmup=1):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def sample_indices(npts, nframes, target_pv=4.8e7, max_frames=64):
    nf = min(nframes, max_frames)
    fidx = np.unique(np.linspace(0, nframes - 1, nf).astype(int))
    n = int(min(npts, max(1000, target_pv // len(fidx))))
    stride = max(1, npts // n)
    return fidx, stride, n


def cpu_vs_gpu_sample(torch, engine, pts, K, spec, wxyz, t, depth_of, masks_of, target_pv=4.8e7, warmup=0, steps=1):
    """Bounded CPU sample of the same workload (every k-th point x evenly spaced frames, about 10-30 s of CPU work) through
    the numpy port, both CPU figures, and the same sample through the CUDA path: must agree bit for bit."""
    from oracle import cpu_baseline as cb
    fidx, stride, n = sample_indices(len(pts), len(t), target_pv)
    sub = np.ascontiguousarray(pts[::stride][:n])
    d = np.stack([depth_of(int(f)) for f in fidx])
    m = np.stack([masks_of(int(f)) for f in fidx])
    wq, tq = wxyz[fidx], t[fidx]
    res, cv, cl, _ = cb.measure(sub, K, spec.width, spec.height, wq, tq, d, m, RADIUS, 0.1, spec.zmax, spec.zmax, C1, warmup, steps)
    tab = engine.FrameTable(K, spec.width, spec.height, wq, tq, spec.zmax)
    pk = engine.pack_frames(torch.as_tensor(d).cuda(), torch.as_tensor(m).cuda())
    gv, gl = engine.fuse_project_vote_resolve(engine.pack_points(sub), tab, pk, None, C1, NCLASSES, RADIUS, 0.1, spec.zmax, THRESHOLD, None)
    ok = bool(np.array_equal(gv.cpu().numpy(), cv) and np.array_equal(gl.cpu().numpy(), cl))
    res.update({"unit": UNIT, "kind": "port", "gpu_matches_bit_exact": ok,
                "sample": f"every {stride}th point ({len(sub)}) x {len(fidx)} evenly spaced frames of the {spec.width}x{spec.height} workload"})
    return res, (sub, wq, tq, d, m, cv, cl)


def workload_config(desc, N, F, W, H, world, shard="contiguous"):
    """The `config` object of the JSON line -- ONE function for the B200 arm and the `--impl reference` arm, so the two lines of a
    driver run carry the same workload description (the residency / cache notes describe the B200 arm's timed region)."""
    if world == 1:
        return {"workload": desc, "points": N, "frames": F, "width": W, "height": H, "nclasses": NCLASSES,
                "frames_layout": "resident in HBM as packed uint32 texels (uint16 depth mm | class << 16, 16x16 tiles) produced by the "
                                 "ingest path's f3d_pack_frames; on-disk contract unchanged", "radius": RADIUS,
                "cache": "inputs larger than L2 (packed frames + votes = %.1f GB per step)" % ((F * H * W * 4 + 4 * N * C1) / 1e9),
                "parallelism": "single GPU"}
    per_gpu = -(-F // world)
    return {"workload": desc, "points": N, "frames_total": F, "frames_per_gpu": per_gpu, "width": W, "height": H,
            "nclasses": NCLASSES, "radius": RADIUS, "shard": shard,
            "frames_layout": "packed uint32 texels resident in HBM (see N = 1)",
            "cache": "inputs larger than L2 (packed frames per GPU %.1f GB + exchange buffers)" % (per_gpu * H * W * 4 / 1e9),
            "parallelism": f"frames sharded over {world} GPUs ({shard}); cloud replicated; vote exchange fused into the kernel: "
                           "slot records written into the owner rank's memory over NVLink (symmetric memory), owner-side merge into "
                           "the dense int32 shard + label resolve; the owners store their int16 labels into every rank's copy from "
                           "inside the merge (peer stores), no collective on the data path"}


def run_reference_arm(args, rank, world):
    """`--impl reference`: the CPU port of the reference path on the host cores, same workload shape, bounded sample.
    Pure numpy: nothing of the CUDA library is imported or mapped here (the sample's depth images are rendered by the
    oracle's own z-buffer splat of the sample points)."""
    if rank != 0:
        return
    scenes = importlib.import_module(PKG_NAME + ".scenes")     # seeded numpy scene generator (no CUDA, no libf3d)
    wl = args.workload or ("C2" if args.gpus == 1 else "C3")    # the same config the B200 arm runs at this --gpus
    cfg_key, desc = WORKLOADS[wl]
    spec = scenes.CONFIGS[cfg_key] if wl != "small" else scenes.scaled_spec("C1", 200_000, 8 if args.gpus == 1 else 16, 320, 240)
    full = spec                                                # the workload the line is about (what the B200 arm runs)
    strong = args.gpus == 1 or wl in ("C3", "small") or args.scaling == "strong"
    if wl == "C3":   # the sample below strides the cloud: do not generate 100 M points for it
        spec = scenes.scaled_spec("C3", npoints=10_000_000)
    from oracle import cpu_baseline as cb
    K = scenes.scaled_intrinsics(spec.width, spec.height)
    wxyz, t = scenes.make_poses(spec)
    pts = scenes.make_cloud(spec)
    fidx, stride, n = sample_indices(len(pts), len(t))
    sub = np.ascontiguousarray(pts[::stride][:n])
    wq, tt = wxyz[fidx], t[fidx]
    m = scenes.block_masks((spec.height, spec.width), len(tt), seed=spec.seed, block=32)
    res, _, _, _ = cb.measure(sub, K, spec.width, spec.height, wq, tt, None, m, RADIUS, 0.1, spec.zmax, spec.zmax, C1,
                              warmup=args.warmup, steps=args.steps)
    sample = (f"every {stride}th point ({len(sub)}) x {len(tt)} evenly spaced frames of the {spec.width}x{spec.height} workload per "
              f"step; depth = oracle z-buffer splat of the sample points")
    val, sec = res["value"], res["seconds"]
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak" if args.gpus == 1 or not strong else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(desc, full.npoints, full.nframes * (1 if strong else args.gpus), full.width, full.height, args.gpus, args.shard),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": res["cores"], "kind": "port", "sample": sample,
                         "host_cores_available": res["host_cores_available"], "frames_stage_seconds": res["frames_stage_seconds"],
                         "segment_seconds": res["segment_seconds"], "single_process": res["single_process"]},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ---------------------------------------------------------------------------------------------------------------------------
# nested configs (N = 1): every BASELINE config in the one driver-run line
# ---------------------------------------------------------------------------------------------------------------------------

def nested_fused_config(torch, mods, key, steps, full_oracle):
    """C1 / C4 / C3 on ONE GPU: resident packed frames, fused step timed with the kernel-only events, parity flag.
    `full_oracle`: the CPU port runs on the WHOLE scene and votes + labels are compared cell for cell; else a bounded CPU
    sample + properties (C3: properties only -- its CPU sample runs in the multi-GPU line, bench.py --gpus N)."""
    engine, scenes, fused = mods
    spec = scenes.CONFIGS[key]
    pts = scenes.make_cloud(spec)
    ids = list(range(spec.nframes))
    fl, K, wxyz, t = make_labeler(fused, scenes, spec, ids, pts)
    big = key == "C3"          # 100 M x 5000: the two-array stacks (41 GB) are not kept next to the packed frames + 54 GB of votes
    depth, masks = build_frames(torch, engine, fl, spec, ids, keep_unpacked=not big)
    N, F, H, W = fl.N, fl.nframes, spec.height, spec.width
    kt = engine.KernelTimer()
    for _ in range(3):
        fl.label()
    fl.stats.zero_()
    ms = timed(torch, lambda: fl.label(timer=kt), steps, warmup=0)
    kms = float(np.mean(kt.ms()))
    st = fl.stats_dict()
    peak, _ = peaks()
    balg = algorithmic_bytes(N, F, H, W, 2, C1)
    out = {"workload": WORKLOADS[key][1], "points": N, "frames": F, "width": W, "height": H, "steps": steps, "ms_per_step": ms,
           "value": float(N) * F / (ms * 1e-3), "unit": UNIT, "kernel_ms": kms, "algorithmic_bytes": balg,
           "roofline_frac": balg / (kms * 1e-3) / 1e9 / peak, "call_frac": balg / (ms * 1e-3) / 1e9 / peak,
           "candidates_per_step": st["candidates"] / steps, "votes_per_step": st["seen"] / steps}
    votes, labels = fl.votes, fl.labels
    props = bool(tensor_sum(votes) * steps == st["seen"] and torch.equal(labels, engine.resolve_labels(votes, NCLASSES, THRESHOLD, None)))
    if full_oracle:
        from oracle import cpu_baseline as cb
        d_np, m_np = depth.cpu().numpy(), masks.cpu().numpy()
        cpu = cb.CpuFusion(pts, K, W, H, wxyz, t, d_np, m_np, RADIUS, 0.1, spec.zmax, spec.zmax, C1)
        sec, cv, cl = cpu.run()
        cpu.close()
        same = bool(np.array_equal(votes.cpu().numpy(), cv) and np.array_equal(labels.cpu().numpy(), cl))
        out.update({"parity": same and props, "parity_how": "CPU port on the WHOLE scene: votes [N,134] and labels compared cell for cell",
                    "cpu_port_seconds": sec, "cpu_port_value": float(N) * F / sec, "cpu_port_cores": cpu.workers})
    elif big:
        ms_lab = timed(torch, lambda: fl.label(want_votes=False), max(2, steps // 2))
        out.update({"parity": props, "parity_how": "sum(votes) == votes cast + fused labels == resolve(votes) at full size; the CPU sample and "
                    "the N-rank == 1-rank comparison of this scene run in the multi-GPU line (bench.py --gpus N)", "labels_only_ms": ms_lab,
                    "note": "single-GPU time of the strong-scaling scene: T(1) for the efficiency of the --gpus 2/4/8 lines"})
    else:
        res, _ = cpu_vs_gpu_sample(torch, engine, pts, K, spec, wxyz, t, lambda f: depth[f].cpu().numpy(), lambda f: masks[f].cpu().numpy(),
                                   target_pv=1.6e7)
        out.update({"parity": bool(res["gpu_matches_bit_exact"]) and props,
                    "parity_how": "CPU port on a sample (" + res["sample"] + ") bit-exact + sum(votes) == votes cast + fused labels == "
                                  "resolve(votes) at full size", "cpu_sample_value": res["value"], "cpu_sample_cores": res["cores"]})
    del fl, depth, masks
    torch.cuda.empty_cache()
    return out


def nested_c5(torch, mods, steps):
    """C5: 200 k boxes through kernel (4): sweep broad phase + exact pair predicate, union-find closure."""
    engine, scenes, _ = mods
    from oracle import f3d_oracle as orc
    lo, hi, group, area = scenes.make_boxes()
    B = len(lo)
    dlo, dhi, dg = engine.as_cuda(lo, torch.float64), engine.as_cuda(hi, torch.float64), engine.as_cuda(group, torch.int32)
    edges = engine.box_pairs_aabb(dlo, dhi, dg)
    E = int(edges.shape[0])
    ms_pairs = timed(torch, lambda: engine.box_pairs_aabb(dlo, dhi, dg, cap=E + 1024), steps)
    ms_brute = timed(torch, lambda: engine.box_pairs_aabb(dlo, dhi, dg, cap=E + 1024, brute_force=True), max(1, steps // 4))
    labels = engine.union_find(B, edges)
    ms_uf = timed(torch, lambda: engine.union_find(B, edges), steps)
    t0 = time.perf_counter()
    oe = orc.box_pairs_aabb(lo, hi, group)
    ol = orc.union_find_labels(B, oe)
    cpu_s = time.perf_counter() - t0
    ge = np.unique(np.sort(edges.cpu().numpy().astype(np.int64), axis=1), axis=0)
    same = bool(np.array_equal(ge, oe) and np.array_equal(labels.cpu().numpy().astype(np.int64), ol))
    peak, _ = peaks()
    balg = 24 * 2 * B + 4 * B + 8 * E    # lo + hi (2 x 24 B per box), group, edges written
    return {"workload": "configs[4] instance merge stress: 200k boxes, closed-interval AABB pairs + union-find", "boxes": B, "edges": E,
            "components": int(len(np.unique(ol))), "pairs_ms": ms_pairs, "pairs_ms_brute_force_kernel": ms_brute, "union_find_ms": ms_uf,
            "ms_per_step": ms_pairs + ms_uf, "value": B / ((ms_pairs + ms_uf) * 1e-3), "unit": "boxes/s", "algorithmic_bytes": balg,
            "roofline_frac": balg / ((ms_pairs + ms_uf) * 1e-3) / 1e9 / peak,
            "roofline_note": "24*B+4*B+8*E bytes (SURVEY 8(d)) over pairs + union-find time; the sweep is latency / compare bound, not HBM bound",
            "parity": same, "parity_how": "edge set and component labels equal the CPU restatement (sort-and-sweep + union-find) on all boxes",
            "cpu_seconds": cpu_s}


def micro_lines(torch, mods, fl, depth, masks, steps):
    """Kernels (2), (3) and level V standalone on the resident C2 scene, each with its own HBM roofline."""
    engine, scenes, _ = mods
    peak, _ = peaks()
    N, F, H, W = fl.N, fl.nframes, fl.H, fl.W
    out = {}
    # kernel (3): label resolve, 4*C1 + 8 bytes per point
    lab = torch.empty(N, dtype=torch.int64, device="cuda")
    ms = timed(torch, lambda: engine.resolve_labels(fl.votes, NCLASSES, THRESHOLD, None, out=lab), steps)
    out["resolve_kernel"] = {"ms": ms, "bytes": (4 * C1 + 8) * N, "roofline_frac": (4 * C1 + 8) * N / (ms * 1e-3) / 1e9 / peak,
                             "unit_bytes": "544 B/point", "parity": bool(torch.equal(lab, fl.labels))}
    # kernel (2): z-buffer splat of 32 frames, 16*N + 4*H*W per frame
    nf = min(32, F)
    zbuf = torch.empty((nf, H * W), dtype=torch.int32, device="cuda")
    dout = torch.empty((nf, H, W), dtype=torch.uint16, device="cuda")
    ms = timed(torch, lambda: engine.zbuffer_splat(fl.points4, fl.table, border=10, frame_begin=0, frame_end=nf, zbuf=zbuf, out=dout), max(2, steps // 4))
    b_survey = nf * (16 * N + 4 * H * W)
    b = 16 * N + nf * H * W * (4 + 4 + 2)     # cloud once for all frames; z-buffer initialised (4) + finalised (4 read, 2 written) per pixel
    views = engine.new_stats()
    engine.zbuffer_splat(fl.points4, fl.table, border=0, frame_begin=0, frame_end=nf, zbuf=zbuf, out=dout, stats=views)
    touched = int((dout.view(torch.int16) != 0).sum())
    seen = engine.stats_dict(views)["seen"]
    out["zbuffer_splat"] = {"ms": ms, "frames": nf, "bytes": b, "roofline_frac": b / (ms * 1e-3) / 1e9 / peak,
                            "unit_bytes": "16*N once (the point-stationary sweep reads the cloud once for all frames) + 10*H*W per frame "
                                          "(uint32 z-buffer initialised, read back, uint16 depth written)",
                            "survey_bytes_16N_plus_4HW_per_frame": b_survey, "parity": bool(torch.equal(engine.zbuffer_splat(
                                fl.points4, fl.table, border=10, frame_begin=0, frame_end=nf, zbuf=zbuf, out=dout), depth[:nf])),
                            "point_views_splatted": seen, "texels_touched": touched, "point_views_per_touched_texel": seen / max(touched, 1),
                            "shared_memory_tile_variant": "not built: a shared-memory z-tile can only save global atomics when several point-views "
                            "of a tile fall on the SAME pixel; here %.3f point-views per touched texel (1.3 mm pixels vs ~1 cm point spacing), so "
                            "there is nothing for an on-chip tile to merge and the frame-tile-stationary variant would add a binning pass of "
                            "16 B per point-view for no saved atomics" % (seen / max(touched, 1))}
    del zbuf, dout
    # level V: uv2pt + mask -> votes (VotingSegmentation.vote), 5*H*W bytes per frame
    uv = engine.fuse_uv2pt(fl.points4, fl.table, fl.frames.slice(0, nf), RADIUS, fl.zmin, fl.zmax, frame_begin=0, frame_end=nf)
    vp = torch.zeros((N, C1), dtype=torch.int32, device="cuda")
    mflat = masks[:nf].reshape(nf, H * W)

    def level_v():
        engine.vote_uv2pt(vp, uv, mflat, 1)
    vp.zero_()
    ms = timed(torch, level_v, 1, warmup=0)          # tags must increase across calls: one timed pass over fresh counters
    engine.vote_finalize(vp)
    ref = engine.fuse_project_vote(fl.points4, fl.table, fl.frames.slice(0, nf), None, C1, RADIUS, fl.zmin, fl.zmax, frame_begin=0, frame_end=nf)
    # uv2pt keeps ONE point per pixel (the highest index), so level V casts at most the level-P votes of the same frames
    out["level_v_vote"] = {"ms": ms, "frames": nf, "ms_per_frame": ms / nf, "bytes": nf * 5 * H * W,
                           "roofline_frac": nf * 5 * H * W / (ms * 1e-3) / 1e9 / peak,
                           "unit_bytes": "5*H*W per frame (int32 uv2pt + uint8 mask) + touched vote sectors; one launch per frame (the "
                                         "per-frame de-dup epoch of votes[idx, cls] += 1, voting.py:98), memset / tag strip not included",
                           "votes": int(vp.sum()), "votes_level_p_same_frames": int(ref.sum())}
    del uv, vp, ref
    return out


# ---------------------------------------------------------------------------------------------------------------------------
def run_single(args, torch, mods):
    engine, scenes, fused = mods
    wl = args.workload or "C2"
    cfg_key, desc = WORKLOADS[wl]
    spec = scenes.CONFIGS[cfg_key] if wl != "small" else scenes.scaled_spec("C1", 200_000, 8, 320, 240)
    pts = scenes.make_cloud(spec)
    ids = list(range(spec.nframes))
    fl, K, wxyz, t = make_labeler(fused, scenes, spec, ids, pts)
    depth, masks = build_frames(torch, engine, fl, spec, ids, keep_unpacked=True)
    N, F, H, W = fl.N, fl.nframes, spec.height, spec.width

    # ---- headline: resident packed frames, one fused launch per step
    for _ in range(args.warmup):
        labels = fl.label()
    torch.cuda.synchronize()
    fl.stats.zero_()
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    kt = engine.KernelTimer()
    torch.cuda.synchronize()
    t_wall0 = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        labels = fl.label(timer=kt)
    ev1.record()
    torch.cuda.synchronize()
    t_wall1 = time.time()
    ms_step = ev0.elapsed_time(ev1) / args.steps
    clocks = sampler.stop(t_wall0, t_wall1)
    pv_step = float(N) * float(F)
    st = fl.stats_dict()
    per_step = {k: v / args.steps for k, v in st.items()}
    launches = 4 * args.steps   # supertile_cull + fuse_kernel + fixup_apply_table + fixup_labels_summary per step
    dev_labels = labels.cpu().numpy()

    kms = float(np.mean(kt.ms()))
    balg = algorithmic_bytes(N, F, H, W, 2, C1)
    peak, how = peaks()
    roof = {"bound": "hbm", "achieved": balg / (kms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": balg / (kms * 1e-3) / 1e9 / peak,
            "traffic": None,
            "kernel": "fuse_kernel<VOTE,U32_T16,HB1> (warp-box cull + project + z-test + packed-texel gather + vote + dense vote write + fused "
                      "label resolve)", "kernel_ms": kms, "kernel_launches_timed": int(len(kt.pairs)),
            "call_ms": ms_step, "call_frac": balg / (ms_step * 1e-3) / 1e9 / peak,
            "call": "f3d_fuse_project_vote_resolve = supertile_cull + fuse_kernel + fixup_apply_table + fixup_labels_summary",
            "algorithmic_bytes": balg, "algorithmic_bytes_formula": "16*N + F*H*W*(1+2) + 64*F + 4*N*134 (SURVEY 8(d), s_d = 2: uint16 depth + "
            "uint8 mask; the packed texel the kernel actually reads is 4 B/pixel, i.e. the figure is conservative)",
            "peak_source": how, "bytes_per_point_view": balg / pv_step}
    prof = profile_figures(wl)
    if prof:
        roof["traffic"] = float(prof["dram_bytes_per_launch"])
        roof["fp32_pipe_pct"] = prof.get("fp32_pipe_pct")
        roof["dram_frac"] = float(prof["dram_bytes_per_launch"]) / (kms * 1e-3) / 1e9 / peak
        roof["profile"] = prof.get("source")
        if prof.get("note"):
            roof["profile_note"] = prof["note"]

    # ---- the same step on the two-array inputs and with the re-pack inside the step
    alt = {}
    n_alt = max(3, args.steps // 5)
    if torch.cuda.mem_get_info()[0] < fl.votes.numel() * 4 * 1.1 + (2 << 30):
        alt["two_array_inputs"] = {"skipped": "a second vote tensor does not fit beside this workload"}
    else:
        v2 = torch.empty_like(fl.votes)
        l2 = torch.empty_like(fl.labels)
        kt2 = engine.KernelTimer()
        ms2 = timed(torch, lambda: engine.fuse_project_vote_resolve(fl.points4, fl.table, depth, masks, C1, NCLASSES, RADIUS, fl.zmin, fl.zmax,
                                                                    THRESHOLD, None, votes=v2, labels=l2, timer=kt2), n_alt)
        same2 = bool(torch.equal(v2, fl.votes) and torch.equal(l2, fl.labels))
        alt["two_array_inputs"] = {"ms_per_step": ms2, "value": pv_step / (ms2 * 1e-3), "kernel_ms": float(np.mean(kt2.ms()[1:])),
                                   "roofline_frac": balg / (float(np.mean(kt2.ms()[1:])) * 1e-3) / 1e9 / peak, "same_result": same2,
                                   "what": "uint16 depth [F,H,W] + uint8 mask [F,H,W] stacks gathered separately (round-1 layout)"}
        del v2, l2

    def repack_step():
        fl.pack(depth, masks)
        fl.label()
    ms3 = timed(torch, repack_step, n_alt)
    alt["with_repack_in_step"] = {"ms_per_step": ms3, "value": pv_step / (ms3 * 1e-3),
                                  "what": "f3d_pack_frames of all frames (reads 3, writes 4 B/pixel) + the fused launch, every step"}

    # ---- end to end through the public API: pinned host buffers in, host labels out (persistent labeler: no allocation in the step)
    e2e = None
    if not args.no_e2e:
        h_depth = torch.empty(depth.shape, dtype=depth.dtype, pin_memory=True)
        h_masks = torch.empty(masks.shape, dtype=masks.dtype, pin_memory=True)
        h_depth.copy_(depth)
        h_masks.copy_(masks)
        h_pts = torch.as_tensor(np.ascontiguousarray(fl.points4.cpu().numpy())).pin_memory()
        torch.cuda.synchronize()
        n_e2e = max(2, min(args.steps, 5))
        out = None
        for i in range(1 + n_e2e):
            if i == 1:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
            out, _ = fused.fuse_labels_from_host(h_pts, K, W, H, wxyz, t, h_depth, h_masks, (0.1, spec.zmax), RADIUS, NCLASSES, THRESHOLD,
                                                 None, chunk_frames=64, labeler=fl)
        torch.cuda.synchronize()
        sec = (time.perf_counter() - t0) / n_e2e
        assert np.array_equal(out, dev_labels), "end-to-end labels differ from the device-resident run"
        h2d = int(h_pts.numel() * 4 + h_depth.numel() * 2 + h_masks.numel())
        e2e = {"value": pv_step / sec, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(N * 8),
               "ms_per_step": sec * 1e3, "steps": n_e2e, "h2d_gb_per_s": h2d / sec / 1e9, "labels_match_device_run": True,
               "api": "fused.fuse_labels_from_host(labeler=persistent): pinned host cloud + uint16 depth + uint8 masks -> device in 64-frame "
                      "chunks on a copy stream, packed on the fly, one fused launch, labels -> pinned host"}
        launches_e2e = (F + 63) // 64 + 4
        e2e["gpu_launches_per_step"] = launches_e2e
        del h_depth, h_masks

    # ---- CPU baseline: the numpy port of the reference path on this box's host cores (bounded sample), both figures
    cpu_b = None
    if not args.no_cpu_baseline:
        cpu_b, _ = cpu_vs_gpu_sample(torch, engine, pts, K, spec, wxyz, t, lambda f: depth[f].cpu().numpy(), lambda f: masks[f].cpu().numpy())

    # ---- the other BASELINE configs and kernels, nested in the same line (bounded: skipped with a note past the time budget)
    configs = {}
    want = [c for c in (args.configs.split(",") if args.configs else []) if c]
    nsteps = max(3, args.steps // 5)
    if wl == "C2" and "micro" in want:
        try:
            configs["micro"] = micro_lines(torch, mods, fl, depth, masks, nsteps)
        except Exception as ex:   # noqa: BLE001
            configs["micro"] = {"error": repr(ex)}
    del depth, masks
    fl.frames = None
    fl.votes = None
    torch.cuda.empty_cache()
    for key in want:
        if key == "micro" or key == wl:
            continue
        if time.time() - T_START > args.budget_s:
            configs[key] = {"skipped": f"time budget of {args.budget_s} s for the whole bench.py run reached"}
            continue
        try:
            if key == "C5":
                configs[key] = nested_c5(torch, mods, nsteps)
            elif key in ("C1", "C4", "C3"):
                configs[key] = nested_fused_config(torch, mods, key, nsteps if key != "C3" else 5, full_oracle=(key == "C1"))
            else:
                configs[key] = {"skipped": "unknown config"}
        except Exception as ex:   # noqa: BLE001
            configs[key] = {"error": repr(ex)}

    line = {
        "metric": METRIC, "value": pv_step / (ms_step * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64",
        "data": "synthetic",
        "config": workload_config(desc, N, F, W, H, 1),
        "roofline": roof, "cpu_baseline": cpu_b, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        "per_step_counts": per_step, "alternatives": alt, "configs": configs, "bench_seconds": time.time() - T_START,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------------
def bind_to_gpu_numa(local_rank):
    """Pin this rank (and, through first touch, its pinned host buffers) to the CPU cores nearest to its GPU."""
    info = {}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [i for i in range(ncpu) if (words[i // 64] >> (i % 64)) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
        info = {"gpu_cpu_affinity": f"{min(cpus)}-{max(cpus)}" if cpus else None, "bound_cores": len(allowed)}
        try:
            info["numa_node"] = pynvml.nvmlDeviceGetNumaNodeId(h)
        except Exception:   # noqa: BLE001
            pass
    except Exception as ex:   # noqa: BLE001
        info = {"error": repr(ex)}
    return info


def run_multi(args, torch, mods, rank, world, local_rank):
    import torch.distributed as dist
    engine, scenes, fused = mods
    parallel = importlib.import_module(PKG_NAME + ".parallel")
    numa = bind_to_gpu_numa(local_rank)
    parallel.init_process_group(local_rank)
    wl = args.workload or "C3"
    cfg_key, desc = WORKLOADS[wl]
    base = scenes.CONFIGS[cfg_key] if wl != "small" else scenes.scaled_spec("C1", 200_000, 16, 320, 240)
    strong = wl in ("C3", "small") or args.scaling == "strong"
    if args.points:
        base = scenes.scaled_spec(cfg_key, npoints=args.points, nframes=args.frames or base.nframes, width=base.width, height=base.height)
    spec = base if strong else scenes.scaled_spec(cfg_key, npoints=base.npoints, nframes=base.nframes * world, width=base.width, height=base.height)
    F_total, H, W = spec.nframes, spec.height, spec.width

    # the cloud: generated once (rank 0), broadcast over NCCL
    N = spec.npoints
    p4 = torch.empty((N, 4), dtype=torch.float32, device="cuda")
    pts = None
    if rank == 0:
        pts = scenes.make_cloud(spec)
        p4[:, :3] = torch.as_tensor(pts).cuda()
        p4[:, 3] = 0
    dist.broadcast(p4, 0)
    torch.cuda.synchronize()

    xchg = parallel.VoteExchange(N, C1, torch.device("cuda", local_rank), compact={"auto": None, "on": True, "off": False}[args.compact])
    stats_total = {}

    def run_shard(mode, steps, warmup):
        ids = parallel.frame_shard_ids(F_total, rank, world, mode)
        fl, K, wxyz, t = make_labeler(fused, scenes, spec, ids, p4)
        build_frames(torch, engine, fl, spec, ids)
        kt = engine.KernelTimer()

        def step(timer=None):
            def fuse(**xargs):
                engine.fuse_project_vote_exchange(fl.points4, fl.table, fl.frames, None, C1, radius=RADIUS, zmin=fl.zmin, zmax=fl.zmax,
                                                  stats=fl.stats, timer=timer, **xargs)
            return xchg.run(fuse, NCLASSES, THRESHOLD, None, check="deferred")
        for _ in range(warmup):
            labels = step()
        xchg.finish()
        torch.cuda.synchronize()
        fl.stats.zero_()
        dist.barrier()
        sampler = ClockSampler(local_rank) if rank == 0 else None
        torch.cuda.synchronize()
        t0w = time.time()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(steps):
            labels = step(kt)
        ev1.record()
        torch.cuda.synchronize()
        t1w = time.time()
        xchg.finish()
        ms_total = ev0.elapsed_time(ev1)
        dist.barrier()
        tt = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        clocks = sampler.stop(t0w, t1w) if sampler else None
        st = fl.stats_dict()
        kall = torch.zeros(world, device="cuda")
        kall[rank] = float(np.mean(kt.ms()))
        dist.all_reduce(kall)
        return {"ms_per_step": float(tt.item()) / steps, "labels": labels.clone(), "fl": fl, "K": K, "wxyz": wxyz, "t": t, "ids": ids,
                "kernel_ms": float(np.mean(kt.ms())), "kernel_ms_per_rank": [round(v, 3) for v in kall.tolist()], "clocks": clocks,
                "stats": {k: v / steps for k, v in st.items()}}

    primary = run_shard(args.shard, args.steps, args.warmup)
    pv_step = float(N) * float(F_total)
    ms_step = primary["ms_per_step"]
    fl = primary["fl"]
    labels = primary["labels"]

    # ---- the whole frame set on every rank: (i) the point-sharded alternative (SURVEY 8(e) "alternative worth measuring": each rank
    # fuses ITS N/G points against ALL frames, no vote exchange at all, labels all-gathered) and (ii) parity (a): rank 0 runs the
    # whole cloud x whole frame set on its ONE GPU (labels only) and the N-rank labels must be identical
    parity = {}
    point_sharded = None
    ok = torch.ones(1, dtype=torch.int32, device="cuda")
    if not (args.no_verify and args.no_point_sharded):
        ids_all = list(range(F_total))
        fl_all, _, _, _ = make_labeler(fused, scenes, spec, ids_all, p4)
        if rank == 0 or not args.no_point_sharded:
            build_frames(torch, engine, fl_all, spec, ids_all)
        if not args.no_point_sharded:
            # points dealt to the ranks in 4096-point super-tiles, round robin: every rank gets the same mix of the scene
            # (a contiguous 1/G of a Morton-sorted building is an octant whose visibility differs from rank to rank)
            ST = 4096
            nsup = -(-N // ST)
            per_ps = -(-nsup // world) * ST
            sup_of = [torch.arange(r, nsup, world, device="cuda") for r in range(world)]
            idx_of = [(su[:, None] * ST + torch.arange(ST, device="cuda")[None, :]).reshape(-1) for su in sup_of]
            idx_of = [ix[ix < N] for ix in idx_of]
            perm = torch.full((world, per_ps), N, dtype=torch.int64, device="cuda")          # slot N = dummy for the padding
            for r in range(world):
                perm[r, :len(idx_of[r])] = idx_of[r]
            mine_idx = idx_of[rank]
            a, b = 0, int(len(mine_idx))
            fl_ps, _, _, _ = make_labeler(fused, scenes, spec, ids_all, p4[mine_idx].contiguous() if b > a else p4[:1])
            fl_ps.frames = fl_all.frames
            lab16 = torch.zeros(per_ps, dtype=torch.int16, device="cuda")
            full16 = torch.zeros(per_ps * world, dtype=torch.int16, device="cuda")
            out16 = torch.zeros(N + 1, dtype=torch.int16, device="cuda")
            kt_ps = engine.KernelTimer()

            def step_ps(timer=None):
                lab = fl_ps.label(timer=timer)
                if b > a:
                    lab16[:b - a].copy_(lab)
                dist.all_gather_into_tensor(full16.view(torch.uint8), lab16.view(torch.uint8))
                out16[perm.view(-1)] = full16                                                # back to cloud order
                return out16
            for _ in range(3):
                step_ps()
            torch.cuda.synchronize()
            dist.barrier()
            n_ps = max(3, args.steps // 2)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(n_ps):
                step_ps(kt_ps)
            ev1.record()
            torch.cuda.synchronize()
            dist.barrier()
            tt = torch.tensor([ev0.elapsed_time(ev1) / n_ps], device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            same_ps = bool(torch.equal(out16[:N].to(torch.int64), labels))
            point_sharded = {"ms_per_step": float(tt.item()), "value": pv_step / (float(tt.item()) * 1e-3), "unit": UNIT, "steps": n_ps,
                             "kernel_ms_rank0": float(np.mean(kt_ps.ms())), "labels_equal_frame_sharded": same_ps,
                             "what": f"every rank holds ALL {F_total} packed frames ({F_total * H * W * 4 / 1e9:.1f} GB) and fuses its N/{world} "
                                     "points (4096-point super-tiles dealt round robin) with the single-GPU fused kernel (dense vote shard + "
                                     "labels), then an int16 label all-gather + un-permute: no vote "
                                     "exchange, but G x the frame ingest; frames-sharded (the headline) is what the north star prescribes"}
            if not same_ps:
                ok[0] = 0
            del fl_ps, lab16, full16, out16, perm, idx_of
        if rank == 0 and not args.no_verify:
            try:
                t0 = time.perf_counter()
                l1 = fl_all.label(want_votes=False)
                torch.cuda.synchronize()
                parity["single_gpu_seconds_first_call"] = time.perf_counter() - t0
                ms1 = timed(torch, lambda: fl_all.label(want_votes=False), 2, warmup=0)
                parity["single_gpu_labels_only_ms"] = ms1
                parity["strong_scaling_efficiency_vs_single_gpu_labels_only"] = ms1 / (world * ms_step)
                parity["strong_scaling_note"] = ("T(1) here is the single-GPU LABELS-ONLY time of the same scene (no 4*N*134-byte vote write), a "
                                                 "conservative T(1); the full single-GPU step is configs.C3 of the --gpus 1 line")
                same = bool(torch.equal(l1, labels))
                parity["multi_gpu_matches_single_gpu"] = same
                parity["single_gpu_how"] = (f"rank 0 runs f3d_fuse_project_vote_resolve (labels only) over the whole cloud and all {F_total} "
                                            f"frames on its one GPU; the {world}-rank labels [N] must be identical")
                if not same:
                    ok[0] = 0
                del l1
            except Exception as ex:   # noqa: BLE001
                parity["single_gpu_error"] = repr(ex)
        del fl_all
        torch.cuda.empty_cache()
    if not args.no_verify:
        # ---- parity (b): a CPU-sized sample through the SAME N-rank exchange path against the numpy port
        try:
            sub_n = min(N, 750_000)
            stride = max(1, N // sub_n)
            fidx = np.unique(np.linspace(0, F_total - 1, min(F_total, 64)).astype(int))
            sub4 = fl.points4[::stride][:sub_n].contiguous()
            sN = int(sub4.shape[0])
            mine = [int(f) for i, f in enumerate(fidx) if i % world == rank]
            spec_s = scenes.scaled_spec(cfg_key, npoints=sN, nframes=spec.nframes, width=W, height=H)
            flr, Kr, wxyz_r, t_r = make_labeler(fused, scenes, spec, mine if mine else [int(fidx[0])], p4)   # full cloud: depth of the real scene
            build_frames(torch, engine, flr, spec, mine if mine else [int(fidx[0])])
            fls = fused.FusedLabeler(sub4, Kr, W, H, wxyz_r[mine if mine else [int(fidx[0])]], t_r[mine if mine else [int(fidx[0])]],
                                     point_range=(0.1, spec.zmax), radius=RADIUS, nclasses=NCLASSES)
            xs = parallel.VoteExchange(sN, C1, torch.device("cuda", local_rank))
            nfr = len(mine)

            def fuse_s(**xargs):
                engine.fuse_project_vote_exchange(fls.points4, fls.table, flr.frames.slice(0, max(nfr, 1)), None, C1, radius=RADIUS, zmin=fls.zmin,
                                                  zmax=fls.zmax, frame_begin=0, frame_end=nfr, **xargs)
            ls = xs.run(fuse_s, NCLASSES, THRESHOLD, None)
            shard_votes = xs.shard[:xs.rows].clone()
            gathered = [torch.empty((xs.per, C1), dtype=torch.int32, device="cuda") for _ in range(world)]
            pad = torch.zeros((xs.per, C1), dtype=torch.int32, device="cuda")
            pad[:xs.rows] = shard_votes
            dist.all_gather(gathered, pad)
            if rank == 0:
                from oracle import cpu_baseline as cb
                # depth / masks of the sample frames come from the ranks' own packed stacks: re-render them here for the CPU side
                fla, _, wq_all, t_all = make_labeler(fused, scenes, spec, [int(f) for f in fidx], p4)
                d_all, m_all = build_frames(torch, engine, fla, spec, [int(f) for f in fidx], keep_unpacked=True)
                res, cv, cl, _ = cb.measure(sub4[:, :3].cpu().numpy(), Kr, W, H, wq_all[fidx], t_all[fidx], d_all.cpu().numpy(), m_all.cpu().numpy(),
                                            RADIUS, 0.1, spec.zmax, spec.zmax, C1)
                gv = torch.cat(gathered)[:sN].cpu().numpy()
                same = bool(np.array_equal(gv, cv) and np.array_equal(ls.cpu().numpy(), cl))
                parity["multi_gpu_matches_cpu_bit_exact"] = same
                parity["cpu_sample"] = f"every {stride}th point ({sN}) x {len(fidx)} evenly spaced frames dealt to the {world} ranks, votes + labels"
                parity["cpu_baseline"] = {"value": res["value"], "unit": UNIT, "cores": res["cores"], "kind": "port", "seconds": res["seconds"],
                                          "single_process": res["single_process"], "host_cores_available": res["host_cores_available"]}
                if not same:
                    ok[0] = 0
                del fla, d_all, m_all
            del xs, flr, fls
            torch.cuda.empty_cache()
        except Exception as ex:   # noqa: BLE001
            parity["cpu_sample_error"] = repr(ex)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        parity["multi_gpu_matches_bit_exact"] = bool(int(ok.item())) and "single_gpu_error" not in parity and "cpu_sample_error" not in parity

    # ---- the other sharding (a real scan arrives contiguous)
    other = None
    if not args.no_other_shard:
        other_mode = "contiguous" if args.shard == "interleaved" else "interleaved"
        del primary["fl"]
        fl_frames_bytes = None
        fl.frames = None
        torch.cuda.empty_cache()
        sec = run_shard(other_mode, max(3, args.steps // 4), 2)
        other = {"shard": other_mode, "ms_per_step": sec["ms_per_step"], "value": pv_step / (sec["ms_per_step"] * 1e-3),
                 "kernel_ms_rank0": sec["kernel_ms"], "kernel_ms_per_rank": sec["kernel_ms_per_rank"],
                 "labels_equal_primary": bool(torch.equal(sec["labels"], labels))}
        fl = sec["fl"]

    # ---- end to end: every rank copies ITS frames from pinned host memory, packs, runs the exchange step, labels land on the host
    e2e = None
    if not args.no_e2e:
        # one host label array for the whole job: a /dev/shm file every rank maps; each rank page-locks and fills the slice
        # of the points it owns (gathering first would only make `world` redundant 0.8 GB device->host copies)
        path = [None]
        if rank == 0:
            try:
                st = os.statvfs("/dev/shm")
                if st.f_bavail * st.f_frsize > N * 8 + (1 << 30):          # a short tmpfs would SIGBUS on first touch
                    path = [f"/dev/shm/f3d_bench_labels_{os.getpid()}"]
                    np.memmap(path[0], dtype=np.int64, mode="w+", shape=(N,)).flush()
            except OSError:
                path = [None]
        dist.broadcast_object_list(path, src=0)
        ok, shared_out = 1, path[0] is not None
        own_a = rank * xchg.per
        try:
            ids = other and parallel.frame_shard_ids(F_total, rank, world, other["shard"]) or primary["ids"]
            dd, mm = build_frames(torch, engine, fl, spec, ids, keep_unpacked=True)
            h_depth = torch.empty(dd.shape, dtype=dd.dtype, pin_memory=True)
            h_masks = torch.empty(mm.shape, dtype=mm.dtype, pin_memory=True)
            h_depth.copy_(dd)
            h_masks.copy_(mm)
            del dd, mm
            h_all = None
            if shared_out:
                h_all = torch.from_numpy(np.memmap(path[0], dtype=np.int64, mode="r+", shape=(N,)))
                h_out = h_all[own_a:own_a + xchg.rows]
                h_out.zero_()                                  # first touch on this rank's NUMA node
                if xchg.rows and int(torch.cuda.cudart().cudaHostRegister(h_out.data_ptr(), h_out.numel() * 8, 0)) != 0:
                    shared_out = False                         # cannot page-lock the mapping
            if not shared_out:                                 # private pinned slice instead
                h_out = torch.empty(xchg.rows, dtype=torch.int64, pin_memory=True)
        except Exception as ex:   # noqa: BLE001
            ok = 0
            print(f"bench.py: rank {rank}: no pinned host memory for the end-to-end arm ({ex})", file=sys.stderr)
        flag = torch.tensor([ok], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()):
            def fuse_e(**xargs):
                engine.fuse_project_vote_exchange(fl.points4, fl.table, fl.frames, None, C1, radius=RADIUS, zmin=fl.zmin, zmax=fl.zmax, **xargs)
            n_e2e = max(2, min(args.steps, 4))
            torch.cuda.synchronize()
            dist.barrier()
            for i in range(1 + n_e2e):
                if i == 1:
                    torch.cuda.synchronize()
                    dist.barrier()
                    t0 = time.perf_counter()
                fl.ingest(h_depth, h_masks, 64)
                lab = xchg.run(fuse_e, NCLASSES, THRESHOLD, None, check="deferred", gather=False)
                h_out.copy_(lab, non_blocking=True)
            torch.cuda.synchronize()
            xchg.finish()
            dist.barrier()
            sec = (time.perf_counter() - t0) / n_e2e
            tt = torch.tensor([sec], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            sec = float(tt.item())
            sh = torch.tensor([int(shared_out)], device="cuda")
            dist.all_reduce(sh, op=dist.ReduceOp.MIN)
            shared_all = bool(sh.item())
            if shared_out and xchg.rows:
                torch.cuda.cudart().cudaHostUnregister(h_out.data_ptr())
            dist.barrier()
            if shared_all:                                      # rank 0 reads the WHOLE shared array (as plain pageable memory)
                same_i = int(torch.equal(h_all.clone().cuda(), labels)) if rank == 0 else 1
            else:
                same_i = int(torch.equal(h_out.clone().cuda(), labels[own_a:own_a + xchg.rows]))
            sm = torch.tensor([same_i], device="cuda")
            dist.all_reduce(sm, op=dist.ReduceOp.MIN)
            same = bool(sm.item())
            del h_out, h_all
            h2d = int(h_depth.numel() * 2 + h_masks.numel())
            tot = torch.tensor([h2d], device="cuda", dtype=torch.int64)
            dist.all_reduce(tot)
            e2e = {"value": pv_step / sec, "unit": UNIT, "h2d_bytes_per_step": int(tot.item()), "d2h_bytes_per_step": int(N * 8),
                   "ms_per_step": sec * 1e3, "steps": n_e2e, "labels_match_device_run": same, "h2d_gb_per_s_per_rank": h2d / sec / 1e9,
                   "api": "per rank: pinned host uint16 depth + uint8 masks of its frames -> FusedLabeler.ingest (copy stream + pack) -> "
                          "parallel.VoteExchange.run(gather=False) -> every rank copies the labels of the points it owns into its slice of ONE "
                          "page-locked host array shared by the ranks (/dev/shm); rank 0 checks the whole array against the device run"}
            e2e["host_labels"] = "one shared page-locked array" if shared_all else "per-rank pinned slices"
            del h_depth, h_masks
        dist.barrier()
        if rank == 0 and path[0] and os.path.exists(path[0]):
            os.unlink(path[0])

    if rank == 0:
        peak, how = peaks()
        kms = primary["kernel_ms"]
        Fr = len(primary["ids"])
        balg = algorithmic_bytes(N, Fr, H, W, 2, C1, rows=xchg.rows)
        roof = {"bound": "hbm", "achieved": balg / (kms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": balg / (kms * 1e-3) / 1e9 / peak,
                "traffic": None, "kernel": "fuse_kernel<VOTE,U32_T16,HB1> in exchange mode on rank 0 (sweep + slot records written to the owners); "
                "the dense shard write happens in slot_merge_kernel", "kernel_ms": kms, "kernel_ms_per_rank": primary["kernel_ms_per_rank"],
                "algorithmic_bytes": balg, "peak_source": how,
                "note": "per-rank bytes: cloud + this rank's frames + this rank's shard of the vote tensor"}
        line = {
            "metric": METRIC, "value": pv_step / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "f32+f64", "data": "synthetic",
            "config": workload_config(desc, N, F_total, W, H, world, args.shard),
            "roofline": roof, "cpu_baseline": parity.get("cpu_baseline"), "e2e": e2e, "gpu_launches": (9 if xchg.compact else 7) * args.steps,
            "gpu_launches_per_step": "supertile_cull(_group), [supertile_compact, dead_directory], fuse_kernel, fixup_apply(_table), exchange_publish, "
                                     "slot_merge, queue_accumulate, queue_relabel (own kernels per rank; torch fills, symmetric-memory barriers and the "
                                     "4-byte NCCL flag reduce not counted)",
            "clocks": primary["clocks"],
            "compacted_launch": bool(xchg.compact), "per_step_counts_rank0": primary["stats"], "parity": parity, "other_shard": other, "point_sharded": point_sharded, "numa": numa,
            "bench_seconds": time.time() - T_START,
        }
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS), help="default: C2 on one GPU, C3 on several")
    ap.add_argument("--configs", default="micro,C1,C5,C4,C3", help="N = 1: other BASELINE configs / kernels nested in the line")
    ap.add_argument("--budget-s", type=float, default=420.0, help="nested configs are skipped once the run is this old")
    ap.add_argument("--compact", default="auto", choices=["auto", "on", "off"],
                    help="N > 1: launch the fused kernel only over the super-tiles a rank's frames can see (auto: 4+ ranks, large cloud)")
    ap.add_argument("--shard", default="contiguous", choices=["interleaved", "contiguous"],
                    help="how the frames of the N-GPU job are dealt to the ranks (same results either way; the other one is timed too)")
    ap.add_argument("--scaling", default=None, choices=["strong", "weak"], help="N > 1 with --workload C2: weak = N x 500 frames")
    ap.add_argument("--points", type=int, default=None, help="N > 1: override the cloud size (experiments)")
    ap.add_argument("--frames", type=int, default=None, help="N > 1: override the total frame count (experiments)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-verify", action="store_true", help="N > 1: skip the single-GPU / CPU parity runs")
    ap.add_argument("--no-other-shard", action="store_true")
    ap.add_argument("--no-point-sharded", action="store_true", help="N > 1: skip the point-sharded alternative (all frames on every rank)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        args.steps = 3 if args.steps is None else args.steps
        args.warmup = 1 if args.warmup is None else args.warmup
        run_reference_arm(args, rank, world)
        return
    args.steps = (100 if world == 1 else 20) if args.steps is None else args.steps
    args.warmup = 3 if args.warmup is None else max(args.warmup, 3)

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    pkg = importlib.import_module(PKG_NAME)
    pkg.load()
    mods = (importlib.import_module(PKG_NAME + ".engine"), importlib.import_module(PKG_NAME + ".scenes"),
            importlib.import_module(PKG_NAME + ".fused"))
    if world == 1:
        run_single(args, torch, mods)
    else:
        run_multi(args, torch, mods, rank, world, local_rank)


if __name__ == "__main__":
    main()
