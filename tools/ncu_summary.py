#!/usr/bin/env python
"""Summarise gpurun_out/launches.csv (ncu launch list) and a .ncu-rep (ncu --set full) into profiles/.

    python tools/ncu_summary.py <tag>          # writes profiles/<tag>_launches.md, profiles/<tag>_fill.md
"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
rep = sys.argv[2] if len(sys.argv) > 2 else os.path.join(OUT, "prof_fill.ncu-rep")
launches = sys.argv[3] if len(sys.argv) > 3 else os.path.join(OUT, "launches.csv")

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum", "sm__cycles_elapsed.max",
]

if os.path.exists(launches):
    lines = [l for l in open(launches) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[row["Metric Unit"]]
        k = row["Kernel Name"][:90]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(ROOT, "profiles", f"{tag}_launches.md"), "w") as f:
        f.write(f"# ncu launch list ({tag}): `ncu --metrics gpu__time_duration.sum --clock-control none`\n\n")
        f.write("Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1]:.3f} | {v[1] / tot:.1%} |\n")
    print("wrote launches summary")

if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(os.path.join(ROOT, "profiles", f"{tag}_fill.md"), "w") as f:
        f.write(f"# ncu --set full ({tag}): {os.path.basename(rep)}\n\n")
        seen = set()
        for r in rows[2:]:
            name = r[idx["Kernel Name"]]
            dur = r[idx["gpu__time_duration.sum"]]
            if (name, dur[:3]) in seen:
                continue
            seen.add((name, dur[:3]))
            f.write(f"## `{name}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in idx:
                    f.write(f"| {k} | {r[idx[k]]} | {units[idx[k]]} |\n")
            f.write("\n")
    print("wrote fill summary")
