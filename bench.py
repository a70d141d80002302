#!/usr/bin/env python
"""Benchmark of the fused label-fusion hot path (BASELINE.json metric: point-view projections / s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload C2|C1|C4|small]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over the whole synthetic scene: kernel (1) fused project + z-test + mask
gather + vote over every frame, then kernel (3) label resolve (N = 1), or per point-chunk fuse -> NCCL
reduce-scatter -> resolve -> all-gather with frames sharded across ranks (N > 1, weak scaling: every rank brings
its own `frames_per_gpu` frames of the same cloud).  A point-view is one (point, frame) pair of the nominal
N_points x N_frames product.  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
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
RADIUS, THRESHOLD = 0.05, 0.5

WORKLOADS = {
    # name: (config key in scenes.CONFIGS, description)
    "C2": ("C2", "configs[1] iOS RTAB-style scan: 10M points x 500 frames 1920x1440 uint16-mm depth + uint8 masks"),
    "C1": ("C1", "configs[0] CPU-reference scene: 1M points x 50 frames 640x480"),
    "C4": ("C4", "configs[3] dense 4K: 20M points x 1000 frames 3840x2160"),
    "small": ("C1", "debug scene: 200k points x 8 frames 320x240"),
}


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


def algorithmic_bytes(N, F, H, W, depth_bytes, C1):
    """SURVEY 8(d) designated figure for kernel (1): every input read once, every output written once."""
    return 16 * N + F * H * W * (1 + depth_bytes) + 64 * F + 4 * N * C1


def build_scene(scenes, engine, fused, spec, frame_lo, frame_hi, torch, frame_ids=None):
    """Cloud + poses on the host (seeded numpy), depth = GPU z-buffer splat of the cloud (kernel 2), block masks on
    the GPU.  Returns the FusedLabeler (cloud + frame table for frames [frame_lo, frame_hi), or the global frame indices
    `frame_ids`) and device depth / masks."""
    K = scenes.scaled_intrinsics(spec.width, spec.height)
    wxyz, t = scenes.make_poses(spec)
    if frame_ids is None:
        frame_ids = list(range(frame_lo, frame_hi))
    frame_lo = frame_ids[0] if len(frame_ids) else 0
    wxyz, t = np.ascontiguousarray(wxyz[frame_ids]), np.ascontiguousarray(t[frame_ids])
    pts = scenes.make_cloud(spec)
    fl = fused.FusedLabeler(pts, K, spec.width, spec.height, wxyz, t, point_range=(0.1, spec.zmax), radius=RADIUS,
                            nclasses=NCLASSES)
    F, H, W = len(t), spec.height, spec.width
    depth = torch.empty((F, H, W), dtype=torch.uint16, device="cuda")
    step = max(1, min(F, (1 << 30) // (H * W * 4)))
    zbuf = torch.empty((step, H * W), dtype=torch.int32, device="cuda")
    for a in range(0, F, step):
        b = min(a + step, F)
        engine.zbuffer_splat(fl.points4, fl.table, border=10, frame_begin=a, frame_end=b, zbuf=zbuf, out=depth[a:b])
    del zbuf
    g = torch.Generator(device="cuda")
    g.manual_seed(spec.seed + 104729 + frame_lo)
    masks = torch.empty((F, H, W), dtype=torch.uint8, device="cuda")
    bh, bw = -(-H // 32), -(-W // 32)
    for a in range(0, F, 32):
        b = min(a + 32, F)
        lab = torch.randint(0, NCLASSES, (b - a, bh, bw), generator=g, device="cuda", dtype=torch.int16)
        lab[torch.rand((b - a, bh, bw), generator=g, device="cuda") < 0.05] = NCLASSES
        m = lab.to(torch.uint8).repeat_interleave(32, dim=1).repeat_interleave(32, dim=2)[:, :H, :W]
        masks[a:b] = m
    torch.cuda.synchronize()
    return fl, pts, K, wxyz, t, depth, masks


def cpu_sample(pts, K, spec, wxyz, t, depth, masks, target_pv=4.8e7):
    """Bounded CPU sample of the same workload: every k-th point x evenly spaced frames (about 10-30 s of CPU work)."""
    F = len(t)
    nf = min(F, 64)
    fidx = np.unique(np.linspace(0, F - 1, nf).astype(int))
    npts = int(min(len(pts), max(1000, target_pv // len(fidx))))
    stride = max(1, len(pts) // npts)
    sub = np.ascontiguousarray(pts[::stride][:npts])
    d = np.stack([depth[int(f)].cpu().numpy() for f in fidx])
    m = np.stack([masks[int(f)].cpu().numpy() for f in fidx])
    desc = f"every {stride}th point ({len(sub)}) x {len(fidx)} evenly spaced frames of the {spec.width}x{spec.height} workload"
    return sub, wxyz[fidx], t[fidx], d, m, desc


def run_reference_arm(args, rank, world):
    """`--impl reference`: the CPU port of the reference path on the host cores, same workload shape, bounded sample.
    Pure numpy: nothing of the CUDA library is imported or mapped here (the sample's depth images are rendered by the
    oracle's own z-buffer splat of the sample points)."""
    if rank != 0:
        return
    scenes = importlib.import_module(PKG_NAME + ".scenes")     # seeded numpy scene generator (no CUDA, no libf3d)
    cfg_key, desc = WORKLOADS[args.workload]
    spec = scenes.CONFIGS[cfg_key] if args.workload != "small" else scenes.scaled_spec("C1", 200_000, 8, 320, 240)
    from oracle import cpu_baseline as cb
    K = scenes.scaled_intrinsics(spec.width, spec.height)
    wxyz, t = scenes.make_poses(spec)
    pts = scenes.make_cloud(spec)
    nf = min(len(t), 64)
    fidx = np.unique(np.linspace(0, len(t) - 1, nf).astype(int))
    npts = int(min(len(pts), max(1000, 4.8e7 // len(fidx))))
    stride = max(1, len(pts) // npts)
    sub = np.ascontiguousarray(pts[::stride][:npts])
    wq, tt = wxyz[fidx], t[fidx]
    m = scenes.block_masks((spec.height, spec.width), len(tt), seed=spec.seed, block=32)
    res, _, _, _ = cb.measure(sub, K, spec.width, spec.height, wq, tt, None, m, RADIUS, 0.1, spec.zmax, spec.zmax, NCLASSES + 1,
                              warmup=args.warmup, steps=args.steps)
    sample = (f"every {stride}th point ({len(sub)}) x {len(tt)} evenly spaced frames of the {spec.width}x{spec.height} workload per "
              f"step; depth = oracle z-buffer splat of the sample points")
    val, sec = res["value"], res["seconds"]
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "points": spec.npoints, "frames": spec.nframes, "width": spec.width,
                   "height": spec.height, "nclasses": NCLASSES},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": res["cores"], "kind": "port", "sample": sample,
                         "host_cores_available": res["host_cores_available"], "frames_stage_seconds": res["frames_stage_seconds"],
                         "segment_seconds": res["segment_seconds"], "single_process": res["single_process"]},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--chunks", type=int, default=4, help="point chunks of the multi-GPU pipeline")
    ap.add_argument("--exchange", default="records", choices=["records", "dense"],
                    help="multi-GPU vote exchange: slot records written by the fused kernel into the owner's memory over "
                         "NVLink (default), or dense packed-uint16 reduce-scatter through NCCL")
    ap.add_argument("--shard", default="interleaved", choices=["interleaved", "contiguous"],
                    help="how the frames of the N-GPU job are dealt to the ranks (same results either way)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        args.steps = 3 if args.steps is None else args.steps
        args.warmup = 1 if args.warmup is None else args.warmup
        run_reference_arm(args, rank, world)
        return
    args.steps = 100 if args.steps is None else args.steps   # 0.25 s timed at C2: enough for several clock samples
    args.warmup = 3 if args.warmup is None else max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    pkg = importlib.import_module(PKG_NAME)
    pkg.load()
    engine = importlib.import_module(PKG_NAME + ".engine")
    scenes = importlib.import_module(PKG_NAME + ".scenes")
    fused = importlib.import_module(PKG_NAME + ".fused")
    parallel = importlib.import_module(PKG_NAME + ".parallel")
    if world > 1:
        parallel.init_process_group(local_rank)

    cfg_key, desc = WORKLOADS[args.workload]
    base = scenes.CONFIGS[cfg_key] if args.workload != "small" else scenes.scaled_spec("C1", 200_000, 8, 320, 240)
    fpg = base.nframes                                   # frames per GPU (weak scaling in frames)
    spec = scenes.scaled_spec(cfg_key, npoints=base.npoints, nframes=fpg * world, width=base.width, height=base.height)
    fl, pts, K, wxyz, t, depth, masks = build_scene(scenes, engine, fused, spec, 0, 0, torch,
                                                    frame_ids=parallel.frame_shard_ids(spec.nframes, rank, world, args.shard))
    N, F, H, W, C1 = fl.N, fl.table.F, spec.height, spec.width, NCLASSES + 1
    stats = fl.stats

    labels_buf = torch.empty(N, dtype=torch.int64, device="cuda")
    ktimer = None   # engine.KernelTimer during the timed region: CUDA events around the fused kernel alone

    def step_single():
        # one launch: kernel (1) with the label resolve (kernel 3's arithmetic) fused into its epilogue
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        votes, labels = engine.fuse_project_vote_resolve(fl.points4, fl.table, depth, masks, C1, NCLASSES, RADIUS, fl.zmin,
                                                         fl.zmax, THRESHOLD, None, votes=fl.votes, labels=labels_buf,
                                                         stats=stats, timer=ktimer)
        e1.record()
        fl.votes = votes
        return labels, (e0, e1), 4   # supertile_cull + fuse_kernel + fixup_apply + fixup_labels

    pipe = xchg = None
    if world > 1 and args.exchange == "dense":
        pipe = parallel.ShardedPipeline(N, C1, args.chunks, torch.device("cuda", local_rank))
    elif world > 1:
        # the record exchange needs peer-mapped symmetric memory (NVLink / NVSwitch P2P); every rank must take the same
        # path, so a failure anywhere falls back to the NCCL reduce-scatter pipeline everywhere
        err = None
        try:
            xchg = parallel.VoteExchange(N, C1, torch.device("cuda", local_rank))
        except Exception as ex:   # noqa: BLE001
            err = ex
        bad = torch.tensor([1 if err is not None else 0], device="cuda")
        dist.all_reduce(bad, op=dist.ReduceOp.MAX)
        if int(bad.item()):
            if rank == 0:
                print(f"bench.py: symmetric-memory vote exchange unavailable ({err}); using the NCCL reduce-scatter pipeline", file=sys.stderr)
            xchg = None
            args.exchange = "dense"
            pipe = parallel.ShardedPipeline(N, C1, args.chunks, torch.device("cuda", local_rank))

    def step_records():
        def fuse(**xargs):
            engine.fuse_project_vote_exchange(fl.points4, fl.table, depth, masks, C1, radius=RADIUS, zmin=fl.zmin, zmax=fl.zmax,
                                              stats=stats, timer=ktimer, **xargs)

        labels = xchg.run(fuse, NCLASSES, THRESHOLD, None)
        return labels, None, 7   # supertile_cull, fuse_kernel, fixup_apply, publish, slot_merge, queue_accumulate, queue_relabel

    def step_multi():
        launches = [0]

        def fuse_into(a, b, out):
            launches[0] += 3   # supertile_cull + fuse_kernel + fixup_apply
            engine.fuse_project_vote(fl.points4[a:b], fl.table, depth, masks, C1, RADIUS, fl.zmin, fl.zmax, votes=out[:b - a],
                                     stats=stats)

        def resolve(v, out):
            launches[0] += 1
            engine.resolve_labels(v, NCLASSES, THRESHOLD, None, out=out)

        labels = pipe.run(fuse_into, resolve)
        return labels, None, launches[0]

    step = step_single if world == 1 else (step_multi if pipe is not None else step_records)
    for _ in range(args.warmup):
        labels, _, _ = step()
    torch.cuda.synchronize()
    stats.zero_()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    torch.cuda.synchronize()
    t_wall0 = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_events, launches = [], 0
    ktimer = engine.KernelTimer()
    ev0.record()
    for _ in range(args.steps):
        labels, kev, nl = step()
        launches += nl
        if kev:
            kernel_events.append(kev)
    ev1.record()
    torch.cuda.synchronize()
    t_wall1 = time.time()
    ms_total = ev0.elapsed_time(ev1)
    if world > 1:
        dist.barrier()
        tt = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    ms_step = ms_total / args.steps
    pv_step = float(N) * float(F) * world
    value = pv_step / (ms_step * 1e-3)
    st = engine.stats_dict(stats)
    per_step = {k: v / args.steps for k, v in st.items()}

    roof = None
    if kernel_events:
        call_ms = float(np.mean([a.elapsed_time(b) for a, b in kernel_events]))   # whole C-ABI call (4 kernels)
        kt = ktimer.ms()
        kms = float(np.mean(kt)) if len(kt) else call_ms
        balg = algorithmic_bytes(N, F, H, W, 2, C1)
        peak, how = peaks()
        ach = balg / (kms * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                "kernel": "fuse_kernel<VOTE,U16MM,HB1> (cull level 2 + project + z-test + gather + vote + dense vote write + "
                          "fused label resolve)", "kernel_ms": kms, "kernel_launches_timed": int(len(kt)),
                "call_ms": call_ms, "call_frac": balg / (call_ms * 1e-3) / 1e9 / peak,
                "call": "f3d_fuse_project_vote_resolve = supertile_cull + fuse_kernel + fixup_apply + fixup_labels",
                "algorithmic_bytes": balg, "peak_source": how, "bytes_per_point_view": balg / (float(N) * float(F))}
        tj = ROOT / "profiles" / "fuse_kernel_traffic.json"
        if tj.exists() and args.workload == "C2":
            try:
                roof["traffic"] = float(json.loads(tj.read_text())["dram_bytes_per_launch"])
            except Exception:
                pass

    if roof is None and xchg is not None:
        kt = ktimer.ms()
        if len(kt):
            kms = float(np.mean(kt))
            balg = 16 * N + F * H * W * 3 + 64 * F + 4 * xchg.rows * C1    # per rank: the dense vote write is this rank's shard
            peak, how = peaks()
            roof = {"bound": "hbm", "achieved": balg / (kms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                    "frac": balg / (kms * 1e-3) / 1e9 / peak, "traffic": None,
                    "kernel": "fuse_kernel<VOTE,U16MM,HB1> in exchange mode on rank 0 (sweep + slot records written to the owners); "
                              "the dense shard write happens in slot_merge_kernel", "kernel_ms": kms,
                    "kernel_launches_timed": int(len(kt)), "algorithmic_bytes": balg, "peak_source": how,
                    "note": "per-rank bytes: cloud + this rank's frames + this rank's shard of the vote tensor"}

    # ---- end to end through the public API: host (pinned) buffers in, host labels out ------------------------------------
    e2e = None
    if not args.no_e2e and world > 1:
        # every rank copies its frames from pinned host memory, the exchange step runs, labels land on the host.  The
        # pinned allocation (4 GB per rank) is the only step that can fail on one rank alone: agree on it first.
        ok = 1
        try:
            h_depth = torch.empty(depth.shape, dtype=depth.dtype, pin_memory=True)
            h_masks = torch.empty(masks.shape, dtype=masks.dtype, pin_memory=True)
            h_depth.copy_(depth)
            h_masks.copy_(masks)
            h_pts = torch.as_tensor(np.ascontiguousarray(fl.points4.cpu().numpy())).pin_memory()
            h_out = torch.empty(N, dtype=torch.int64, pin_memory=True)
        except Exception as ex:   # noqa: BLE001
            ok = 0
            print(f"bench.py: rank {rank}: no pinned host memory for the end-to-end arm ({ex})", file=sys.stderr)
        flag = torch.tensor([ok], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()):
            ref_labels = labels.clone()
            n_e2e = max(2, min(args.steps, 5))
            torch.cuda.synchronize()
            dist.barrier()
            for i in range(1 + n_e2e):
                if i == 1:
                    torch.cuda.synchronize()
                    dist.barrier()
                    t0 = time.perf_counter()
                fl.points4.copy_(h_pts, non_blocking=True)
                depth.copy_(h_depth, non_blocking=True)
                masks.copy_(h_masks, non_blocking=True)
                lab = step()[0]
                h_out.copy_(lab, non_blocking=True)
            torch.cuda.synchronize()
            dist.barrier()
            sec = (time.perf_counter() - t0) / n_e2e
            tt = torch.tensor([sec], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            sec = float(tt.item())
            same = bool(torch.equal(torch.as_tensor(h_out.numpy()).cuda(), ref_labels))
            h2d = int(h_pts.numel() * 4 + h_depth.numel() * 2 + h_masks.numel())
            e2e = {"value": pv_step / sec, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": int(N * 8) * world,
                   "ms_per_step": sec * 1e3, "steps": n_e2e, "labels_match_device_run": same,
                   "api": "per rank: pinned host cloud + frames -> device, parallel.VoteExchange.run, labels -> pinned host"}
            del h_depth, h_masks
    if not args.no_e2e and world == 1:
        h_depth = torch.empty(depth.shape, dtype=depth.dtype, pin_memory=True)
        h_masks = torch.empty(masks.shape, dtype=masks.dtype, pin_memory=True)
        h_depth.copy_(depth)
        h_masks.copy_(masks)
        h_pts = torch.as_tensor(pts).pin_memory()
        dev_labels = labels.cpu().numpy()
        del fl.votes
        fl.votes = None
        torch.cuda.synchronize()
        n_e2e = max(2, min(args.steps, 5))
        out = None
        for i in range(1 + n_e2e):
            if i == 1:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
            out, fl2 = fused.fuse_labels_from_host(h_pts, K, W, H, wxyz, t, h_depth, h_masks, (0.1, spec.zmax), RADIUS,
                                                   NCLASSES, THRESHOLD, None)
            del fl2
        torch.cuda.synchronize()
        sec = (time.perf_counter() - t0) / n_e2e
        assert np.array_equal(out, dev_labels), "end-to-end labels differ from the device-resident run"
        h2d = int(h_pts.numel() * 4 + h_depth.numel() * 2 + h_masks.numel() + len(t) * 7 * 8)
        e2e = {"value": pv_step / sec, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(N * 8),
               "ms_per_step": sec * 1e3, "steps": n_e2e,
               "api": "fused.fuse_labels_from_host (pinned host frames streamed in 64-frame chunks)"}
        del h_depth, h_masks

    # ---- CPU baseline: the numpy port of the reference path on this box's host cores (bounded sample) ------------------
    cpu_b = None
    if not args.no_cpu_baseline and world == 1 and rank == 0:
        from oracle import cpu_baseline as cb
        sub, wq, tq, d, m, sdesc = cpu_sample(pts, K, spec, wxyz, t, depth, masks)
        cpu = cb.CpuFusion(sub, K, W, H, wq, tq, d, m, RADIUS, 0.1, spec.zmax, spec.zmax, C1)
        sec, cv, cl = cpu.run()
        cpu.close()
        # the same sample on the GPU must agree bit for bit (the oracle stays the checker, never the product)
        tab = engine.FrameTable(K, W, H, wq, tq, spec.zmax)
        gv = engine.fuse_project_vote(engine.pack_points(sub), tab, torch.as_tensor(d).cuda(), torch.as_tensor(m).cuda(), C1,
                                      RADIUS, 0.1, spec.zmax)
        gl = engine.resolve_labels(gv, NCLASSES, THRESHOLD, None)
        ok = bool(np.array_equal(gv.cpu().numpy(), cv) and np.array_equal(gl.cpu().numpy(), cl))
        cpu_b = {"value": len(sub) * len(tq) / sec, "unit": UNIT, "cores": cpu.workers, "kind": "port", "sample": sdesc,
                 "seconds": sec, "host_cores_available": cpu.cores, "gpu_matches_bit_exact": ok}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64",
            "data": "synthetic",
            "config": {"workload": desc, "points": N, "frames_per_gpu": F, "frames_total": F * world, "width": W, "height": H,
                       "nclasses": NCLASSES, "depth": "uint16 mm", "radius": RADIUS, "cache": "inputs larger than L2 "
                       "(depth+masks+votes = %.1f GB per GPU)" % ((F * H * W * 3 + 4 * N * C1) / 1e9),
                       "parallelism": "single GPU" if world == 1 else f"frames sharded over {world} GPUs ({args.shard}), "
                       + ("vote exchange fused into the kernel: slot records written into the owner's memory over NVLink, "
                          "owner-side merge into the dense shard + labels, all-gather of labels" if args.exchange == "records" else
                          f"{args.chunks}-chunk pipeline: fuse -> NCCL reduce-scatter of packed uint16 votes -> resolve -> "
                          "all-gather")},
            "roofline": roof, "cpu_baseline": cpu_b, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "per_step_counts": per_step,
        }
        print(json.dumps(line))
    if world > 1:
        if xchg is not None:
            xchg.check_overflow()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
