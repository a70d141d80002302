"""Golden g10: the call signatures of the reference functions the package mirrors, read from the reference SOURCE with `ast`
(nothing is imported or executed), so that `tests/test_signatures_golden.py` can check the drop-in surface without the reference.

    python tests/golden/make_golden_signatures.py        # build container only: needs /root/reference
"""
import ast
import json
from pathlib import Path

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent / "g10_signatures.json"
# reference file -> the mirror module inside the package that restates (part of) it
FILES = {
    "get3DSeg.py": "get3DSeg",
    "Fusion3DSeg/camera_utils.py": "Fusion3DSeg.camera_utils",
    "Fusion3DSeg/intersections.py": "Fusion3DSeg.intersections",
    "Fusion3DSeg/fusion.py": "Fusion3DSeg.fusion",
    "Fusion3DSeg/process3D.py": "Fusion3DSeg.process3D",
    "Fusion3DSeg/merge_intersecting_bb.py": "Fusion3DSeg.merge_intersecting_bb",
    "Fusion3DSeg/segUtils/voting.py": "Fusion3DSeg.segUtils.voting",
    "Fusion3DSeg/segUtils/cv.py": "Fusion3DSeg.segUtils.cv",
    "RTAB_utils/ios_rtab.py": "RTAB_utils.ios_rtab",
    "RTAB_utils/spatQuad.py": "RTAB_utils.spatQuad",
}


def describe(fn: ast.FunctionDef):
    a = fn.args
    pos = [x.arg for x in a.posonlyargs + a.args]
    defaults = [None] * (len(pos) - len(a.defaults)) + [ast.unparse(d) for d in a.defaults]
    return {"params": pos, "defaults": defaults, "vararg": a.vararg.arg if a.vararg else None,
            "kwonly": [x.arg for x in a.kwonlyargs], "kwarg": a.kwarg.arg if a.kwarg else None, "line": fn.lineno,
            "decorators": [ast.unparse(d) for d in fn.decorator_list]}


def main():
    out = {}
    for rel, mirror in FILES.items():
        tree = ast.parse((REF / rel).read_text())
        entry = {"mirror": mirror, "functions": {}, "classes": {}}
        for node in tree.body:
            if isinstance(node, ast.FunctionDef):
                entry["functions"][node.name] = describe(node)
            elif isinstance(node, ast.ClassDef):
                entry["classes"][node.name] = {m.name: describe(m) for m in node.body if isinstance(m, ast.FunctionDef)}
        out[rel] = entry
    OUT.write_text(json.dumps(out, indent=1, sort_keys=True) + "\n")
    print("wrote", OUT, sum(len(e["functions"]) + sum(len(c) for c in e["classes"].values()) for e in out.values()), "signatures")


if __name__ == "__main__":
    main()
