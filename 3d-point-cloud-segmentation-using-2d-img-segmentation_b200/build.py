"""Compile csrc/*.cu into libf3d.so for sm_100a with plain nvcc (no torch dependency in the library).

    python <package>/build.py [--force]

The shared object is built in-tree (git-ignored, shipped to the GPU box with the snapshot).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libf3d.so"
SOURCES = ["c_api.cu", "frame_setup.cu", "fuse_vote.cu", "fuse_aux.cu", "vote_resolve.cu", "vote_exchange.cu", "box_merge.cu", "geometry.cu"]
COMPILE_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-fno-fast-math,-ffp-contract=off", "-Xptxas", "-v",
]
LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: libf3d has no CPU fallback and cannot be built without the CUDA toolkit")


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + \
        [PKG.parent / "include" / "f3d.h", Path(__file__)]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False, extra_defs=(), out: Path | None = None) -> Path:
    """Compile every translation unit in parallel (nvcc -c, objects under build/obj*/), then link the shared object.
    `extra_defs` / `out`: kernel-variant experiments (tools/) only."""
    lib = LIB if out is None else Path(out)
    if not force and out is None and not needs_build():
        return lib
    from concurrent.futures import ThreadPoolExecutor
    objdir = PKG.parent / "build" / ("obj" if out is None else "obj_" + lib.stem)
    objdir.mkdir(parents=True, exist_ok=True)
    nvcc = nvcc_path()

    def compile_one(src):
        obj = objdir / (Path(src).stem + ".o")
        cmd = [nvcc, *COMPILE_FLAGS, *extra_defs, "-c", "-o", str(obj), str(CSRC / src)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, res.returncode, res.stdout + res.stderr

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    log = "".join(f"==== {src}\n{out_}" for src, _, _, out_ in results)
    bad = [src for src, _, rc, _ in results if rc != 0]
    if not bad:
        res = subprocess.run([nvcc, *LINK_FLAGS, "-o", str(lib), *[str(o) for _, o, _, _ in results]], capture_output=True, text=True)
        log += "==== link\n" + res.stdout + res.stderr
        if res.returncode != 0:
            bad = ["link"]
    (PKG / "csrc" / "build.log" if out is None else lib.with_suffix(".log")).write_text(log)
    if bad:
        raise RuntimeError(f"nvcc failed ({', '.join(bad)}):\n" + log[-6000:])
    if verbose:
        print(log)
    return lib


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print("built", LIB)
