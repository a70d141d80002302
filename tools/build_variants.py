"""Kernel-variant experiments: builds libf3d with different -D settings into build/variants/ (git-ignored, shipped by gpurun).
usage: python tools/build_variants.py name1="-DFUSE_BLOCK=128 -DFUSE_MINB8=5" name2="..." """
import importlib, re, sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
b = importlib.import_module("3d-point-cloud-segmentation-using-2d-img-segmentation_b200.build")
out = ROOT / "build" / "variants"
out.mkdir(parents=True, exist_ok=True)
specs = [a.split("=", 1) for a in sys.argv[1:]]
def one(spec):
    name, defs = spec
    lib = out / f"libf3d_{name}.so"
    b.build(force=True, extra_defs=defs.split(), out=lib)
    log = lib.with_suffix(".log").read_text()
    rows = []
    for m in re.finditer(r"Compiling entry function '(_Z11fuse_kernel\w+)' for 'sm_100a'\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers", log):
        if "ELi0ELi3ELi1ELb0" in m.group(1) or "ELi0ELi0ELi1ELb0" in m.group(1):
            rows.append(f"  fmt{m.group(1)[24]}: regs {m.group(5)} stack {m.group(2)} spill st/ld {m.group(3)}/{m.group(4)}")
    return name, defs, rows
with ThreadPoolExecutor(max_workers=4) as ex:
    for name, defs, rows in ex.map(one, specs):
        print(f"== {name}: {defs}")
        print("\n".join(rows))
