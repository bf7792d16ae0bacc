"""Summaries of the round-2 ncu captures (scripts/profile_all.sh) from the CSV exports in
gpurun_out/: one block of counters per kernel -> profiles/r02_kernels_summary.txt, the launch
list -> profiles/r02_launch_list_summary.csv, the integrator's source-level breakdown ->
profiles/r02_integrate_kernel_summary.txt and its DRAM traffic -> profiles/r02_integrate_traffic.json."""
import collections
import csv
import json
import math
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "gpurun_out")
WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe active %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sass__inst_executed_register_spilling", "spill instructions executed"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
]
kernels = ["integrate_kernel", "gbs_restore_kernel", "derivatives_grain_kernel", "derivatives_gbm_kernel",
           "voigt_average_kernel", "mindex_pairs_reflect_live_kernel", "mindex_pairs_reflect_kernel",
           "mindex_quats_kernel", "pathlines_kernel",
           "resample_kernel", "pack_soa_kernel", "fp64_peak_kernel"]
lines = ["ncu --set full --clock-control none, one launch per kernel, bench size (1250 particles x 2 phases x",
         "5000 grains; M-index: 64 textures of 5000 grains); commands: scripts/profile_all.sh, scripts/final_run.sh",
         "(mindex_pairs_reflect_kernel is the all-49-dots cross-check kernel, DREX_MINDEX_FULL=1)", ""]
traffic = None
for k in kernels:
    path = os.path.join(OUT, f"r02_{k}_raw.csv")
    if not os.path.exists(path):
        lines += [f"== {k}: no capture", ""]
        continue
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    lines.append(f"== drex::{k}" if k != "fp64_peak_kernel" else f"== {k} (the roofline denominator)")
    for key, label in WANT:
        if key in d:
            lines.append(f"   {label:32s} {d[key][1]} {d[key][0]}")
    stalls = {h.split("issue_stalled_")[1]: float(v) for h, (u, v) in d.items()
              if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h and v not in ("", "n/a")}
    tot = sum(stalls.values())
    if tot > 0:
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:6]
        lines.append("   top stall reasons (share of samples): " + ", ".join(f"{n} {100 * v / tot:.0f}%" for n, v in top))
    lines.append("")
    if k == "integrate_kernel":
        def val(key):
            u, v = d[key]
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
            return float(v) * scale
        traffic = {"workload": "corner_flow", "particles_per_gpu": 1250, "grains": 5000,
                   "dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                   "launch_ms_under_ncu": float(d["gpu__time_duration.sum"][1]),
                   "source": "gpurun_out/r02_integrate_kernel_raw.csv (ncu --set full, launch 7 of scripts/profile_kernels.py)"}
open(os.path.join(HERE, "r02_kernels_summary.txt"), "w").write("\n".join(lines))
if traffic:
    json.dump(traffic, open(os.path.join(HERE, "r02_integrate_traffic.json"), "w"), indent=1)

# integrator: instruction mix and where the samples fall
src = os.path.join(OUT, "r02_integrate_kernel_src.csv")
if os.path.exists(src):
    text = subprocess.run([sys.executable, os.path.join(HERE, "analyze_ncu.py"),
                           os.path.join(OUT, "r02_integrate_kernel_raw.csv"), src],
                          capture_output=True, text=True).stdout
    text = "\n".join(l for l in text.splitlines() if not l.startswith("instr "))
    rows = list(csv.reader(open(src)))
    hdr = rows[1]
    ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    bins = collections.defaultdict(lambda: [0, 0, 0])
    tot_n = tot_s = 0
    for r in rows[2:]:
        try:
            n, s = int(r[ie]), int(r[isamp])
        except (ValueError, IndexError):
            continue
        tot_n += n
        tot_s += s
        if n:
            b = round(math.log10(n) * 8) / 8
            bins[b][0] += 1
            bins[b][1] += n
            bins[b][2] += s
    text += "\n\ninstructions grouped by how often they ran in this launch (392500 chunks of 32 grains; ~2.4 M olivine and\n"
    text += "~1.0 M enstatite warp-level right-hand sides; ~0.75 M olivine steps):\n"
    for b in sorted(bins):
        st, n, s = bins[b]
        if s > 0.004 * tot_s or n > 0.004 * tot_n:
            text += f"  executed ~{10 ** b:9.0f} times: {st:5d} static instructions, {100 * n / tot_n:5.1f}% of executed, {100 * s / tot_s:5.1f}% of stall samples\n"
    open(os.path.join(HERE, "r02_integrate_kernel_summary.txt"), "w").write(text)

# launch list of the bench command itself (scripts/final_run.sh); else of scripts/profile_kernels.py
raw = os.path.join(OUT, "r02_bench_launches_raw.csv")
if not os.path.exists(raw):
    raw = os.path.join(OUT, "r02_launches_raw.csv")
if os.path.exists(raw):
    out = subprocess.run([sys.executable, os.path.join(HERE, "summarize_launches.py"), raw],
                         capture_output=True, text=True).stdout
    open(os.path.join(HERE, "r02_launch_list_summary.csv"), "w").write(out)
    rows = [r for r in csv.reader(open(raw)) if len(r) > 10 and r[0].isdigit()]
    with open(os.path.join(HERE, "r02_launches_raw.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel", "block", "grid", "duration_ns"])
        for r in rows:
            w.writerow([r[0], r[4].split("(")[0][:60], r[7], r[8], r[-1]])
print(open(os.path.join(HERE, "r02_kernels_summary.txt")).read())
