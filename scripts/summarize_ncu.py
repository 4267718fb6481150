#!/usr/bin/env python
"""Turn gpurun_out/{launches_TAG.csv, prof_TAG.ncu-rep, bench_TAG.json} into tracked summaries under profiles/.

  python scripts/summarize_ncu.py TAG [round-label]
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
label = sys.argv[2] if len(sys.argv) > 2 else tag
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)
go = os.path.join(ROOT, "gpurun_out")

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "lts__t_sectors_op_read.sum", "lts__t_bytes.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]

lines = [f"# ncu summary {label}", ""]
bench = os.path.join(go, f"bench_{tag}.json")
if os.path.exists(bench) and os.path.getsize(bench):
    lines += ["## bench line (plain run, no profiler)", "", "```json", open(bench).read().strip(), "```", ""]

launch_csv = os.path.join(go, f"launches_{tag}.csv")
if os.path.exists(launch_csv):
    rows = [r for r in csv.reader(open(launch_csv)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    per = {}
    order = []
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki]
        if name not in per:
            per[name] = []
            order.append(name)
        per[name].append(v)
    unit = rows[1][hdr.index("Metric Unit")] if len(rows) > 1 else "ns"
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(unit, 1e-6)
    tot = sum(sum(v) for v in per.values()) * scale
    lines += [f"## launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, cold-cache + serialised: compare shares)", "",
              "| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for name in sorted(order, key=lambda n: -sum(per[n])):
        t = sum(per[name]) * scale
        lines.append(f"| `{name[:90]}` | {len(per[name])} | {t:.3f} | {100 * t / tot:.1f}% |")
    lines.append("")
    with open(os.path.join(out_dir, f"{label}_launches.csv"), "w") as f:
        f.write(open(launch_csv).read())

rep = os.path.join(go, f"prof_{tag}.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    lines += ["## `ncu --set full --clock-control none --import-source on` (per captured launch)", ""]
    # DRAM traffic per launch of every captured kernel -> profiles/traffic.json (bench.py reports it as roofline.traffic)
    tpath = os.path.join(out_dir, "traffic.json")
    traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
    unit_scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[2:]:
        try:
            ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            b = float(r[ir]) * unit_scale.get(units[ir], 1.0) + float(r[iw]) * unit_scale.get(units[iw], 1.0)
            name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
            key = f"{label}:{name}"
            traffic[key] = max(traffic.get(key, 0.0), b)      # several captured launches of one kernel: the largest (the motif launch)
        except (ValueError, IndexError):
            pass
    json.dump(traffic, open(tpath, "w"), indent=1, sort_keys=True)
    for r in rows[2:]:
        lines += [f"### {r[hdr.index('Kernel Name')]}  grid {r[hdr.index('launch__grid_size')]} x block {r[hdr.index('launch__block_size')]}", "",
                  "| metric | value | unit |", "|---|---|---|"]
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                lines.append(f"| {k} | {r[i]} | {units[i]} |")
        lines.append("")
    det = subprocess.run(["ncu", "-i", rep, "--page", "details", "--launch-count", "1"], capture_output=True, text=True).stdout
    keep = [l for l in det.splitlines() if l.strip()]
    with open(os.path.join(out_dir, f"{label}_details.txt"), "w") as f:
        f.write("\n".join(keep) + "\n")
    lines += [f"Full `--page details` text of the first captured launch: `profiles/{label}_details.txt`.", ""]

with open(os.path.join(out_dir, f"{label}_summary.md"), "w") as f:
    f.write("\n".join(lines) + "\n")
print("\n".join(lines[:60]))
