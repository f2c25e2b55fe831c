#!/usr/bin/env python
"""Regenerate profiles/ncu_traffic.json from `ncu --set full` captures of THIS build (run here, no GPU needed):

    python tools/ncu_traffic.py <tag> key=report.ncu-rep[:units_per_launch[:kernel-regex,...]] ...

Each key sums dram__bytes_read + dram__bytes_write and smsp__inst_executed over the launches of the listed kernels
(all launches in the report when no regex is given) and records the units (pairs, bases) one launch group covers.
bench.py scales the figures to its own launch size; nothing here is a timing."""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
out = {"_comment": f"dram__bytes_read.sum + dram__bytes_write.sum and smsp__inst_executed.sum per launch group from the ncu "
                   f"--set full captures of round tag {tag} (same build as the bench numbers); written by tools/ncu_traffic.py"}


def num(v):
    return float(v.replace(",", "") or 0)


UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
for spec in sys.argv[2:]:
    key, rest = spec.split("=", 1)
    parts = rest.split(":")
    rep = parts[0]
    units = int(parts[1]) if len(parts) > 1 and parts[1] else 262144
    pats = [re.compile(p) for p in parts[2].split(",")] if len(parts) > 2 and parts[2] else None
    if rep.endswith(".csv"):  # `ncu -i x.ncu-rep --page raw --csv` exported on the GPU box (reports are too big to bring back)
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, un = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    ir, iw, ii, it = (hdr.index(k) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
                                             "gpu__time_duration.sum"))
    dram = inst = 0.0
    kernels = []
    for r in rows[2:]:
        if pats and not any(p.search(r[ik]) for p in pats):
            continue
        dram += num(r[ir]) * UNIT[un[ir]] + num(r[iw]) * UNIT[un[iw]]
        inst += num(r[ii])
        kernels.append(f"{r[ik][:60]} ({r[it]} {un[it]})")
    out[key] = {"dram_bytes": dram, "inst_executed": inst, "units_per_launch": units, "source": os.path.basename(rep),
                "kernels": kernels}
with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))
