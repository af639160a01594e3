#!/usr/bin/env python
"""Turn gpurun_out/ ncu artefacts into the tracked summaries under profiles/.

  python profiles/summarize.py <tag> [launches.csv] [prof.ncu-rep]

writes profiles/<tag>_launches.csv (copy), profiles/<tag>_launch_summary.md (per-kernel share
of a step), profiles/<tag>_k_cost_ncu.txt (key counters of the dominant kernel) and
profiles/k_cost_traffic.json (DRAM bytes per launch, read by bench.py's roofline.traffic).
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_blocks",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__thread_inst_executed_pred_on_per_inst_executed.ratio",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
    "sm__cycles_elapsed.max",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
]


def launches(tag, path):
    shutil.copy(path, os.path.join(HERE, f"{tag}_launches.csv"))
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    start = next(i for i, r in enumerate(rows) if r[0] == "ID")
    idx = {h: i for i, h in enumerate(rows[start])}
    seq = []
    for r in rows[start + 1:]:
        try:
            name = r[idx["Kernel Name"]].split("(")[0]
            v = float(r[idx["Metric Value"]].replace(",", ""))
            unit = r[idx["Metric Unit"]]
        except Exception:
            continue
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        seq.append((name, r[idx["Grid Size"]], v))
    # one device-resident step = from a full-size k_undistort to the following k_chain_apply
    big = max((g for n, g, _ in seq if "k_cost" in n and "track" not in n), key=lambda g: int(g.strip("()").split(",")[0]))
    steps, cur = [], None
    for n, g, v in seq:
        if "k_undistort" in n:
            cur = []
        if cur is not None:
            cur.append((n, g, v))
            if "k_chain_apply" in n:
                if any(gg == big for nn, gg, _ in cur if "k_cost" in nn):
                    steps.append(cur)
                cur = None
    agg = collections.OrderedDict()
    for st in steps:
        for n, g, v in st:
            agg.setdefault((n, g), []).append(v)
    tot = sum(sum(v) / len(v) for v in agg.values())
    with open(os.path.join(HERE, f"{tag}_launch_summary.md"), "w") as f:
        f.write(f"# {tag}: per-kernel time of one device-resident step (ncu gpu__time_duration, serialised, cold cache)\n\n")
        f.write(f"{len(steps)} full steps found in the launch list; mean per launch.\n\n| kernel | grid | us | share |\n|---|---|---:|---:|\n")
        for (n, g), v in agg.items():
            m = sum(v) / len(v)
            f.write(f"| {n} | {g} | {m:.1f} | {m / tot:.3f} |\n")
        f.write(f"| **total** | | {tot:.1f} | 1.000 |\n")
    print(open(os.path.join(HERE, f"{tag}_launch_summary.md")).read())


def kernel(tag, rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    out = []
    traffic = None
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    seen = set()
    for r in rows[2:]:
        name = r[ix["Kernel Name"]]
        key = (name, r[ix["launch__grid_size"]])
        if key in seen:
            continue
        seen.add(key)
        out.append(f"== {name}  grid {r[ix['launch__grid_size']]}")
        for k in KEYS:
            if k in ix:
                out.append(f"{k:85s} {units[ix[k]]:16s} {r[ix[k]]}")
        out.append("")
        if "k_cost" in name and "track" not in name and traffic is None:
            rd = float(r[ix["dram__bytes_read.sum"]]) * scale[units[ix["dram__bytes_read.sum"]]]
            wr = float(r[ix["dram__bytes_write.sum"]]) * scale[units[ix["dram__bytes_write.sum"]]]
            traffic = {"dram_bytes_per_launch": rd + wr, "read": rd, "write": wr, "source": f"profiles/{tag}_k_cost_ncu.txt",
                       "grid": r[ix["launch__grid_size"]]}
    open(os.path.join(HERE, f"{tag}_k_cost_ncu.txt"), "w").write("\n".join(out) + "\n")
    if traffic:
        json.dump(traffic, open(os.path.join(HERE, "k_cost_traffic.json"), "w"), indent=1)
    print("\n".join(out))


if __name__ == "__main__":
    tag = sys.argv[1]
    lp = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "launches.csv")
    rp = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", "prof_k_cost.ncu-rep")
    if os.path.exists(lp):
        launches(tag, lp)
    if os.path.exists(rp):
        kernel(tag, rp)
