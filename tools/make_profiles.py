"""Build the tracked round-2 summaries under profiles/ from the scratch captures in gpurun_out/ (run in the build container).

    python tools/make_profiles.py

Inputs (all produced through `gpurun`; commands in profiles/README.md): prof_fuse_r2.ncu-rep (ncu --set full of the fused kernel
inside bench.py), launches_r2.csv (ncu launch list of bench.py), bench_*.json lines, micro-benchmark logs."""
import csv, json, subprocess, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402  (kernel_source_hash)

G, P = ROOT / "gpurun_out", ROOT / "profiles"
P.mkdir(exist_ok=True)


def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2]


rep = G / "prof_fuse_r2.ncu-rep"
if rep.exists():
    hdr, units, vals = raw_metrics(rep)
    m = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
            "lts__t_requests_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sector_hit_rate.pct",
            "l1tex__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
            "launch__block_size", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
    want += sorted(h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio"))
    txt = ["# ncu --set full --clock-control none, kernel fuse_kernel<VOTE,U32_T16,HB1,noaudit> (128-point tiles, 96 regs, 5 CTAs/SM), bench.py C2 "
           "workload (packed resident frames), launch 7 of that instantiation (after the warm-ups)",
           "# command: ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:fuse_kernelILi0ELi3 -s 6 -c 1 "
           "python bench.py --steps 5 --configs \"\" --no-cpu-baseline --no-e2e",
           f"# kernel source hash {bench.kernel_source_hash()} (csrc/fuse_kernel.cuh + f3d_common.cuh)"]
    for w in want:
        if w in m:
            txt.append(f"{w:90s} {m[w][0]:16s} {m[w][1]}")
    (P / "r2_fuse_kernel_ncu_full.txt").write_text("\n".join(txt) + "\n")

    def gb(name):
        u, v = m[name]
        v = float(v.replace(",", ""))
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
    traffic = {"workload": "C2", "dram_bytes_read": gb("dram__bytes_read.sum"), "dram_bytes_write": gb("dram__bytes_write.sum"),
               "dram_bytes_per_launch": gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum"),
               "fp32_pipe_pct": float(m["sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"][1]),
               "fma_pipe_pct_of_active": float(m["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"][1]),
               "issue_active_pct": float(m["smsp__issue_active.avg.pct_of_peak_sustained_active"][1]),
               "kernel_ms_under_ncu": float(m["gpu__time_duration.sum"][1]) * {"ms": 1.0, "us": 1e-3, "ns": 1e-6}[m["gpu__time_duration.sum"][0]],
               "kernel_source_hash": bench.kernel_source_hash(), "source": "profiles/r2_fuse_kernel_ncu_full.txt (ncu --set full, one launch)"}
    (P / "fuse_kernel_traffic.json").write_text(json.dumps(traffic, indent=1) + "\n")
    print("traffic", traffic)

    # hot source lines: warp-stall samples and executed instructions per CUDA source line
    out = subprocess.run(["ncu", "-i", str(rep), "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur = hdr2 = None
    agg = {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] in ("Function Name",):
            continue
        if r[0] == "Line No":
            hdr2 = r
            continue
        if hdr2 is None or not r[0] or r[2] != "-":
            continue
        d = dict(zip(hdr2[4:], r[4:]))
        try:
            agg[(cur, int(r[0]))] = (int(d["# Samples"]), int(d["Instructions Executed"]), r[1].strip()[:110], d)
        except ValueError:
            continue
    ts, ti = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
    lines = [f"# fuse_kernel<VOTE,U32_T16,HB1>: top source lines by warp-stall samples ({ts} samples, {ti} warp instructions per launch); same capture as "
             "r2_fuse_kernel_ncu_full.txt"]
    for (f, l), (s, i, src, d) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
        st = sorted(((k, int(v)) for k, v in d.items() if k.startswith("stall_") and "(" not in k and v.isdigit() and int(v) > 0), key=lambda kv: -kv[1])[:2]
        lines.append(f"{f}:{l:5d} samples {100 * s / ts:5.2f}% inst {100 * i / ti:5.2f}% {st} | {src}")
    (P / "r2_fuse_kernel_hot_lines.txt").write_text("\n".join(lines) + "\n")

src = G / "launches_r2.csv"
if src.exists():
    rows = [l for l in src.read_text().splitlines() if l.startswith('"')]
    (P / "r2_launches_bench_c2.csv").write_text("\n".join(rows) + "\n")

for name, dst in (("gather_layouts.log", "r2_gather_layouts.txt"), ("gather_fetch.log", "r2_gather_fetch_granularity.txt")):
    if (G / name).exists():
        (P / dst).write_text((G / name).read_text())
for name in ("gather_fetch_ncu_default.csv",):
    if (G / name).exists():
        rows = [r for r in csv.reader(open(G / name)) if len(r) > 10]
        h = rows[0]
        ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
        d = {}
        for r in rows[1:]:
            d.setdefault((int(r[ii]), r[ki][:28]), {})[r[mi]] = r[vi]
        names = {0: "__ldg (ld.global.nc)", 1: "plain ld.global", 2: "ld.global.nc.L1::no_allocate", 3: "ld.global.cg", 4: "ld.global.nc.L2::64B",
                 5: "ld.global.cs", 6: "ld.global.lu"}
        out = ["# tools/micro/gather_fetch.cu under ncu: 80.0 M random 4-byte gathers (45x38 px window per warp, 16x16-tile texel layout)",
               "# load flavour                      time us   L2 requests   L2 sectors read   DRAM sectors read   DRAM sectors / gather"]
        for (i, k), v in sorted(d.items()):
            if i % 2 == 0:
                continue
            out.append(f"{names[i // 2]:34s} {float(v['gpu__time_duration.sum']) / 1e3:8.1f} {v['lts__t_requests_srcunit_tex_op_read.sum']:>13s} "
                       f"{v['lts__t_sectors_srcunit_tex_op_read.sum']:>17s} {v['dram__sectors_read.sum']:>19s} {float(v['dram__sectors_read.sum']) / 8.0e7:10.2f}")
        (P / "r2_gather_fetch_granularity_ncu.txt").write_text("\n".join(out) + "\n")
lines = []
for n in ("bench_d.json", "bench_n2.json", "bench_n2_compact.json", "bench_n4.json", "bench_n8.json", "bench_ref.json"):
    f = G / n
    if f.exists():
        ls = [l for l in f.read_text().splitlines() if l.startswith("{")]
        if ls:
            lines.append(ls[-1])
if lines:
    (P / "r2_bench_lines.jsonl").write_text("\n".join(lines) + "\n")
print("profiles written:", sorted(p.name for p in P.iterdir()))
