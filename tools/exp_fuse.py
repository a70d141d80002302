"""Timing experiments on the fused kernel (GPU box only): phase switches via the debug bits of `flags`."""
import importlib, sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench
PKG = bench.PKG_NAME
engine = importlib.import_module(PKG + ".engine"); scenes = importlib.import_module(PKG + ".scenes"); fused = importlib.import_module(PKG + ".fused")
from importlib import import_module
_lib = import_module(PKG + "._lib")
wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
spec = scenes.CONFIGS[wl]
fl, pts, K, wxyz, t, depth, masks = bench.build_scene(scenes, engine, fused, spec, 0, spec.nframes, torch)
N, C1 = fl.N, 134
votes = torch.empty((N, C1), dtype=torch.int32, device="cuda"); labels = torch.empty(N, dtype=torch.int64, device="cuda")
lib = _lib.load()
ws = engine.workspace(N, fl.points4.device)
def run(flags, want_votes=True, want_labels=True, reps=5):
    def call():
        if want_labels:
            rc = lib.f3d_fuse_project_vote_resolve(fl.points4.data_ptr(), N, fl.table.table.data_ptr(), 0, fl.table.F, depth.data_ptr(), 0, masks.data_ptr(), fl.table.H, fl.table.W, fl.table.K.ctypes.data, 0.05, 0.1, spec.zmax, votes.data_ptr() if want_votes else None, C1, 0.5, None, 0, 133, labels.data_ptr(), ws.data_ptr(), ws.numel(), None, flags, torch.cuda.current_stream().cuda_stream)
        else:
            rc = lib.f3d_fuse_project_vote(fl.points4.data_ptr(), N, fl.table.table.data_ptr(), 0, fl.table.F, depth.data_ptr(), 0, masks.data_ptr(), fl.table.H, fl.table.W, fl.table.K.ctypes.data, 0.05, 0.1, spec.zmax, votes.data_ptr(), C1, 0, ws.data_ptr(), ws.numel(), None, flags, torch.cuda.current_stream().cuda_stream)
        assert rc == 0, lib.f3d_last_error()
    for _ in range(2): call()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): call()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
mode = sys.argv[2] if len(sys.argv) > 2 else ""
if mode in ("quick", "traffic"):
    # every kernel variant under build/variants/ (tools/build_variants.sh) on the same resident scene; "traffic" does
    # three launches per variant (run it under ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum)
    import ctypes, glob, os
    ref = None
    reps = 10 if mode == "quick" else 1
    for path in ["default", "default+F3D_NO_SUPERTILE", "default+F3D_FIXUP_STAGING", "default+F3D_NO_SUMMARY"] + sorted(glob.glob(str(ROOT / "build" / "variants" / "*.so"))) + ["default+F3D_HIST16"]:
        os.environ.pop("F3D_HIST16", None); os.environ.pop("F3D_NO_SUPERTILE", None); os.environ.pop("F3D_FIXUP_STAGING", None); os.environ.pop("F3D_NO_SUMMARY", None)
        if "NO_SUPERTILE" in path: os.environ["F3D_NO_SUPERTILE"] = "1"
        if "FIXUP_STAGING" in path: os.environ["F3D_FIXUP_STAGING"] = "1"
        if "NO_SUMMARY" in path: os.environ["F3D_NO_SUMMARY"] = "1"
        if path.startswith("default"):
            lib = _lib.load()
            if "HIST16" in path: os.environ["F3D_HIST16"] = "1"
        else:
            lib = ctypes.CDLL(path)
            for name, (res, args) in _lib.SIGNATURES.items():
                fn = getattr(lib, name); fn.restype = res; fn.argtypes = args
        t = run(0, True, True, reps=reps)
        tv = run(0, True, False, reps=reps) if mode == "quick" else 0.0
        sig = (int(votes.sum()), int(labels.sum()))
        ref = ref or sig
        print(f"{os.path.basename(path):28s} votes+labels {t:7.3f} ms   votes only {tv:7.3f} ms   {'same' if sig == ref else 'DIFFERENT ' + str(sig)}", flush=True)
    sys.exit(0)
for name, fl_, wv, wl_ in [("full votes+labels", 0, True, True), ("votes only", 0, True, False), ("labels only (no vote write)", 0, False, True),
                      ("classify only (no gathers)", 0x400, True, True), ("gathers, no phase 3", 0x800, True, True), ("no deferred fp64", 0x1000, True, True), ("cull only, no candidates", 0x100, True, True),
                      ("no cull, no candidates", 0x200, True, True), ("no cull, labels only", 0x200, False, True)]:
    print(f"{name:34s} {run(fl_, wv, wl_):8.3f} ms", flush=True)
# plain memset of the vote tensor for reference
torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): votes.zero_()
e1.record(); torch.cuda.synchronize(); print(f"{'torch zero_ of votes (5.36 GB)':34s} {e0.elapsed_time(e1)/5:8.3f} ms")
# exchange mode with purely local buffers (G = 1): isolates the cost of the record logic from NVLink stores
parallel = importlib.import_module(PKG + ".parallel")
nreg, nsub, _, nlev = engine.exchange_constants()
per = parallel.shard_points(N, 1); sub_rows = max(256, -(-(per // 32 * 40) // nreg)); sub_cap = max(512, -(-per // nsub))
queue = torch.empty(nsub * sub_cap, dtype=torch.int64, device="cuda"); counts = torch.zeros(nsub, dtype=torch.int32, device="cuda")
cursors = torch.zeros(nreg + nsub, dtype=torch.int32, device="cuda"); ovf = torch.zeros(1, dtype=torch.int32, device="cuda")
slots = torch.empty(nreg * sub_rows * 32, dtype=torch.uint16, device="cuda"); dirs = torch.empty(per // 32 * nlev, dtype=torch.int64, device="cuda")
P = lambda t_: np.array([t_.data_ptr()], dtype=np.uint64)
def timed(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
def records_local():
    cursors.zero_()
    engine.fuse_project_vote_exchange(fl.points4, fl.table, depth, masks, C1, 1, per, P(slots), P(dirs), P(queue), sub_rows, sub_cap, cursors, ovf, 0.05, 0.1, spec.zmax)
print(f"{'exchange records, local (G=1)':34s} {timed(records_local):8.3f} ms  rows={int(cursors[:nreg].sum())} queue={int(cursors[nreg:].sum())} overflow={int(ovf.item())}")
engine.exchange_publish(cursors, P(counts), 0, sub_cap)
shard = torch.empty((per, C1), dtype=torch.int32, device="cuda"); lab = torch.empty(per, dtype=torch.int64, device="cuda")
print(f"{'exchange merge (all points)':34s} {timed(lambda: engine.exchange_merge(slots, dirs, 1, sub_rows, per, N, C1, 133, 0.5, None, votes=shard, labels=lab)):8.3f} ms")
print(f"{'exchange queue apply':34s} {timed(lambda: engine.exchange_queue_apply(queue, counts, 1, sub_cap, shard, N, 133, lab, 0.5, None)):8.3f} ms")
