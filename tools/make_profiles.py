"""Build the tracked summaries under profiles/ from the scratch captures in gpurun_out/ (run in the build container).

    python tools/make_profiles.py

Inputs (all produced through `gpurun`, see profiles/README.md for the commands): bench_*.json, launches_r1f.csv,
prof_fuse_r1f.ncu-rep, phases_r1f.log, exp_gather.log, multi2.log, multi8.log, multi8c.log."""
import csv, json, subprocess, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
G, P = ROOT / "gpurun_out", ROOT / "profiles"
P.mkdir(exist_ok=True)


def last_json(name):
    f = G / name
    if not f.exists():
        return None
    lines = [l for l in f.read_text().splitlines() if l.startswith("{")]
    return lines[-1] if lines else None


lines = [l for l in (last_json(n) for n in ("bench_c2.json", "bench_ref.json", "bench_c1.json", "bench_n2.json", "bench_n4.json",
                                            "bench_n8.json")) if l]
(P / "r1_bench_lines.jsonl").write_text("\n".join(lines) + "\n")
BENCH_LINES = lines

# launch list: keep the csv rows only (drop ncu banner lines)
src = G / "launches_r1f.csv"
if src.exists():
    rows = [l for l in src.read_text().splitlines() if l.startswith('"')]
    (P / "r1_launches_bench_c2.csv").write_text("\n".join(rows) + "\n")

rep = G / "prof_fuse_r1f.ncu-rep"
if rep.exists():
    out = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    want = ["dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
            "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
            "launch__block_size", "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
            "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum"]
    txt = ["# ncu --set full --clock-control none, kernel fuse_kernel<VOTE,U16MM,HB1> (80 regs, 3 CTAs/SM), bench.py C2 workload, launch 4 (after 3 warm-ups)",
           "# command: ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:fuse_kernelILi0ELi0 -s 3 -c 1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"]
    d = {}
    for i, h in enumerate(hdr):
        if h in want or ("issue_stalled" in h and "per_issue_active" in h):
            txt.append(f"{h:90s} {units[i]:18s} {vals[i]}")
            d[h] = vals[i]
    (P / "r1_fuse_kernel_ncu_full.txt").write_text("\n".join(txt) + "\n")
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    def bytes_of(key):
        i = hdr.index(key)
        return float(vals[i].replace(",", "")) * scale[units[i]]
    rd, wr = bytes_of("dram__bytes_read.sum"), bytes_of("dram__bytes_write.sum")
    (P / "fuse_kernel_traffic.json").write_text(json.dumps({
        "workload": "C2", "kernel": "fuse_kernel<VOTE,U16MM,HB1>", "dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd,
        "dram_bytes_write": wr, "source": "profiles/r1_fuse_kernel_ncu_full.txt"}, indent=1) + "\n")
    # the ncu capture and the C2 bench line come from the same gpurun call: give the stored line this capture's traffic
    fixed = []
    for l in BENCH_LINES:
        d = json.loads(l)
        if d.get("n_gpus") == 1 and d.get("roofline") and "configs[1]" in d["config"]["workload"] and d.get("impl") != "reference":
            d["roofline"]["traffic"] = rd + wr
        fixed.append(json.dumps(d))
    (P / "r1_bench_lines.jsonl").write_text("\n".join(fixed) + "\n")

for a, b in (("phases_r1f.log", "r1_fuse_phases.txt"), ("exp_gather.log", "r1_gather_ceiling.txt")):
    if (G / a).exists():
        (P / b).write_text((G / a).read_text())

multi = []
for name, what in (("multi2.log", "2 GPUs, interleaved frames"), ("multi8.log", "8 GPUs, interleaved frames"),
                   ("multi8c.log", "8 GPUs, contiguous frames (before the last merge / label-gather changes)")):
    f = G / name
    if f.exists():
        keep = [l for l in f.read_text().splitlines() if l.startswith(("records:", "single-GPU", "dense", "labels agree"))]
        multi += [f"# tools/exp_multi.py records -- {what}"] + keep + [""]
if multi:
    (P / "r1_multi_gpu_step_breakdown.txt").write_text("\n".join(multi))
print("profiles/ updated:", sorted(p.name for p in P.iterdir()))
