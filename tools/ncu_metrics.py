#!/usr/bin/env python
"""Key metrics of the first kernel in a .ncu-rep (ncu --set full): python tools/ncu_metrics.py gpurun_out/x.ncu-rep [kernel-index]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = open(rep).read() if rep.endswith(".csv") else \
    subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2 + which]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "sm__cycles_elapsed.max",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.avg.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
for h, u, v in zip(hdr, units, vals):
    stall = "issue_stalled" in h and h.endswith("_per_warp_active.pct")
    if h in KEYS or (stall and float(v.replace(",", "") or 0) > 1.0):
        print(f"{h:92s} {v} {u}")
