"""Small-scene bisect (GPU box): every library under build/variants + the in-tree one, formats 0 / 2 / 3, against the oracle."""
import ctypes, glob, importlib, os, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench
PKG = bench.PKG_NAME
engine = importlib.import_module(PKG + ".engine"); scenes = importlib.import_module(PKG + ".scenes"); fused = importlib.import_module(PKG + ".fused")
_lib = importlib.import_module(PKG + "._lib")
from oracle import f3d_oracle as orc
spec = scenes.scaled_spec("C1", npoints=40000, nframes=6, width=320, height=240, seed=5)
K = scenes.scaled_intrinsics(spec.width, spec.height); wxyz, t = scenes.make_poses(spec); pts = scenes.make_cloud(spec)
masks = scenes.block_masks((spec.height, spec.width), spec.nframes, seed=5, block=8)
fl = fused.FusedLabeler(pts, K, spec.width, spec.height, wxyz, t, point_range=(0.1, 4.0), radius=0.05)
depth = fl.render_depth(border=2); m = torch.as_tensor(masks).cuda()
ovotes = orc.fuse_project_vote(pts, K, spec.width, spec.height, wxyz, t, depth.cpu().numpy(), masks, 134, 0, 0.05, 0.1, 4.0, 4.0)
print("oracle votes", ovotes.sum())
N, C1 = fl.N, 134
ws = engine.workspace(N, fl.points4.device)
packed = {2: engine.pack_frames(depth, m, 2), 3: engine.pack_frames(depth, m, 3)}
# check the pack itself
tex2 = packed[2].texels.cpu().numpy().reshape(spec.nframes, spec.height, spec.width)
print("pack linear ok", np.array_equal(tex2 & 0xffff, depth.cpu().numpy()), np.array_equal((tex2 >> 16) & 0xff, masks))
for path in sorted(glob.glob(str(ROOT / "build" / "variants" / "*.so"))) + [str(_lib.LIB_PATH)]:
    if path.endswith("_r1.so"): continue
    lib = ctypes.CDLL(path); fn = lib.f3d_fuse_project_vote; fn.restype, fn.argtypes = _lib.SIGNATURES["f3d_fuse_project_vote"]
    for fmt in (0, 2, 3):
        for use_ws in (True, False):
            votes = torch.full((N, C1), -7, dtype=torch.int32, device="cuda")
            rc = fn(fl.points4.data_ptr(), N, fl.table.table.data_ptr(), 0, fl.table.F, depth.data_ptr() if fmt == 0 else packed[fmt].texels.data_ptr(), fmt,
                    m.data_ptr() if fmt == 0 else None, spec.height, spec.width, fl.table.K.ctypes.data, 0.05, 0.1, 4.0, votes.data_ptr(), C1, 0,
                    ws.data_ptr() if use_ws else None, ws.numel() if use_ws else 0, None, 0, torch.cuda.current_stream().cuda_stream)
            v = votes.cpu().numpy()
            bad = np.argwhere(v != ovotes)
            print(f"{os.path.basename(path):22s} fmt {fmt} ws {int(use_ws)} rc {rc} votes {v.sum():8d} mismatching cells {len(bad)}", bad[:4].tolist() if len(bad) else "", flush=True)
