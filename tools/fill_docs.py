#!/usr/bin/env python3
"""Fill the measured numbers of DESIGN.md / README.md from the bench lines committed under profiles/.

Every number in those two files that comes from a bench run sits between an HTML-comment pair
`<!--KEY-->value<!--/-->`; this script rewrites the values (idempotent, so it is re-run after each new bench).

    python tools/fill_docs.py [--one profiles/r02q_bench_1gpu.json] [--multi 'profiles/r02q_bench_[248]gpu.json']
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def f0(x):
    return f"{x:,.0f}".replace(",", " ")


def values(d):
    ow = d["other_workloads"]
    so, e, c4, c5 = d["score_only"], d["e2e"], ow["c4_global_10kx10k_traceback"], ow["c5_gsw_paired_2x150"]
    rag = ow["c3_ragged_300-500x100-150"]
    ext = ow["gsw_extend_1M_pairs_175x100"]
    v = {
        "C3DEV": f0(d["value"]), "C3BYTES": f0(d["traceback_byte_inputs"]["value"]), "C3E2E": f0(e["value"]),
        "C3PIN": f0(e["variants"]["pinned_bytes"]["value"]), "C3PTB": f0(e["variants"]["pageable_twobit"]["value"]),
        "C3NTB": f0(e["variants"]["pinned_twobit"]["value"]), "C3FRAC": f"{d['roofline']['frac']:.3f}",
        "C3UNP": f0(e["variants"]["pageable_bytes_staged_unpacked"]["value"]),
        "C2UNP": f0(so["e2e"]["pageable_bytes_staged_unpacked"]["value"]),
        "C2DEV": f0(so["value"]), "C2BYTES": f0(so["byte_inputs"]["value"]),
        "C2E2E": f0(so["e2e"]["pageable_bytes"]["value"]), "C2PIN": f0(so["e2e"]["pinned_bytes"]["value"]),
        "C2PTB": f0(so["e2e"]["pageable_twobit"]["value"]), "C2NTB": f0(so["e2e"]["pinned_twobit"]["value"]),
        "C2ISSUE": f"{so.get('issue_roofline', {}).get('frac', float('nan')):.2f}",
        "C4DEV": f0(c4["value"]), "C4E2E": f0(c4["e2e"]["value"]), "C4FRAC": f"{c4['roofline']['frac']:.3f}",
        "RAGT": f0(rag["traceback"]["value"]), "RAGS": f0(rag["score_only"]["value"]),
        "C1": f0(ow["c1_global_1000x150_traceback"]["value"]), "CONST": f0(ow["const_gap_500x150_traceback"]["value"]),
        "PACK": f0(ow["twobit_pack_2Gbase"]["value"]),
        "SEEDS": f"{ow['gsw_seeds_1M_reads_150bp']['value']:.0f}",
        "EXTL": f"{ext['left']['Mpairs_per_s']:.0f}", "EXTR": f"{ext['right']['Mpairs_per_s']:.0f}",
        "C5": f"{c5['value']:.2f}", "C5S": f"{c5['seconds']:.1f}",
        "C5CPU": f"{c5.get('cpu_baseline', {}).get('value', float('nan')):.4f}",
        "CLOCK": f"{d['clocks'].get('sm_mhz') or 0:.0f}",
    }
    cb = d.get("cpu_baseline")
    if cb:
        v["CPU"], v["CPUCORES"] = f"{cb['value']:.2f}", str(cb["cores"])
    return v


def scaling_table(lines):
    rows = ["| GPUs | C3 GCUPS (traceback + CIGAR, gather in the step) | vs 1 GPU | gather ms (GB/s per rank) | C2 GCUPS | C4 GCUPS | C5 Mpairs/s |",
            "|---:|---:|---:|---|---:|---:|---:|"]
    base = None
    for d in sorted(lines, key=lambda x: x["n_gpus"]):
        n = d["n_gpus"]
        if n == 1:
            base = d["value"]
        g = d.get("gather")
        ow = d["other_workloads"]
        c4 = ow.get("c4_global_10kx10k_traceback", {})
        c5 = ow.get("c5_gsw_paired_2x150", {})
        c4s = f0(c4["value"]) + ("" if c4.get("pairs_per_gpu") == 12500 else f" ({c4.get('pairs_per_gpu')} pairs/GPU)") if c4 else ""
        c5s = f"{c5['value']:.2f}" + ("" if c5.get("pairs_total") == 10_000_000 else f" ({c5.get('pairs_total'):.0e} pairs)") if c5 else ""
        rows.append(f"| {n} | {f0(d['value'])} | {d['value'] / base:.2f}× |" if base else f"| {n} | {f0(d['value'])} | |")
        rows[-1] += (f" {g['ms']:.2f} ({g['GBps_per_rank']:.0f}) |" if g else " — |") + f" {f0(d['score_only']['value'])} | {c4s} | {c5s} |"
    return "\n".join(rows)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--one", default=os.path.join(ROOT, "profiles/r02q_bench_1gpu.json"))
    ap.add_argument("--multi", default=os.path.join(ROOT, "profiles/r02q_bench_[248]gpu.json"))
    a = ap.parse_args()
    one = json.load(open(a.one))
    v = values(one)
    v["SCALING"] = "\n" + scaling_table([one] + [json.load(open(p)) for p in sorted(glob.glob(a.multi))]) + "\n"
    for name in ("DESIGN.md", "README.md"):
        p = os.path.join(ROOT, name)
        s = open(p).read()
        missing = set()

        def sub(m):
            k = m.group(1)
            if k not in v:
                missing.add(k)
                return m.group(0)
            return f"<!--{k}-->{v[k]}<!--/-->"
        s2 = re.sub(r"<!--([A-Z0-9]+)-->.*?<!--/-->", sub, s, flags=re.S)
        open(p, "w").write(s2)
        print(name, "updated" if s2 != s else "unchanged", ("missing: " + ",".join(sorted(missing))) if missing else "")


if __name__ == "__main__":
    main()
