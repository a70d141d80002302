"""Micro-benchmark: how fast can a B200 do random 2-byte / 1-byte gathers from multi-GB arrays (sector-granular DRAM reads)?"""
import torch, time
dev = "cuda"
n_src = 1_382_400_000  # 500 x 1440 x 1920 pixels
depth = torch.zeros(n_src, dtype=torch.int16, device=dev)
mask = torch.zeros(n_src, dtype=torch.uint8, device=dev)
for n_idx in (65_000_000, 130_000_000):
    for name, gen in (("uniform random", lambda: torch.randint(0, n_src, (n_idx,), device=dev)),
                      ("windowed (32 pts in a 64x64 px window of one frame)", None)):
        if gen is None:
            nwin = n_idx // 32
            base_f = torch.randint(0, 500, (nwin,), device=dev) * (1440 * 1920)
            bx = torch.randint(0, 1920 - 64, (nwin,), device=dev); by = torch.randint(0, 1440 - 64, (nwin,), device=dev)
            ox = torch.randint(0, 64, (nwin, 32), device=dev); oy = torch.randint(0, 64, (nwin, 32), device=dev)
            idx = (base_f[:, None] + (by[:, None] + oy) * 1920 + bx[:, None] + ox).reshape(-1)
        else:
            idx = gen()
        for arr, nm in ((depth, "u16"), (mask, "u8")):
            for _ in range(2): out = arr[idx]
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3): out = arr[idx]
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            print(f"{name:52s} n={n_idx/1e6:5.0f}M {nm:3s}: {ms:7.3f} ms  -> {n_idx/ms/1e6:7.1f} G gathers/s", flush=True)
        del idx
