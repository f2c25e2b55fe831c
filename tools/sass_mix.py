#!/usr/bin/env python
"""Instruction-mix of the loops of a kernel from cuobjdump SASS: ALU-pipe vs FMA-pipe vs other.

    python tools/sass_mix.py <lib.so> <mangled-name-substring> [min_len]
Pipe classes follow the microbenchmark in profiles/r01_microbench_pipes.txt.
"""
import collections
import re
import subprocess
import sys

lib, pat = sys.argv[1], sys.argv[2]
min_len = int(sys.argv[3]) if len(sys.argv) > 3 else 30
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
ALU = ("VIMNMX", "VIADDMNMX", "LOP3", "PRMT", "SHF", "IADD3", "VIADD", "ISETP", "SEL", "IMNMX", "LEA", "POPC", "FLO",
       "BREV", "PLOP3", "FMNMX", "ICMP", "IABS", "SGXT", "BMSK", "P2R", "R2P", "FSETP", "VABSDIFF")
FMA = ("IMAD", "FFMA", "FADD", "FMUL")
for f in funcs:
    name = f.split("\n")[0].strip()
    if pat not in name:
        continue
    ins = []
    for m in re.finditer(r"/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)(.*?);", f):
        ins.append((int(m.group(1), 16), m.group(3), m.group(4)))
    addr_idx = {a: i for i, (a, _, _) in enumerate(ins)}
    print(f"== {name}: {len(ins)} instructions")
    loops = []
    for i, (a, op, rest) in enumerate(ins):
        if op.startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)", rest)
            if m:
                tgt = int(m.group(1), 16)
                if tgt <= a and tgt in addr_idx:
                    loops.append((addr_idx[tgt], i))
    for lo, hi in loops:
        n = hi - lo + 1
        if n < min_len:
            continue
        # skip loops that strictly contain another listed loop of decent size (outer loops)
        if any(l2 > lo and h2 < hi and (h2 - l2) >= min_len for l2, h2 in loops):
            kind = "outer"
        else:
            kind = "inner"
        hist = collections.Counter()
        alu = fma = other = 0
        for a, op, rest in ins[lo:hi + 1]:
            base = op.split(".")[0]
            hist[base] += 1
            if base.startswith(ALU):
                alu += 1
            elif base.startswith(FMA):
                fma += 1
            else:
                other += 1
        print(f"  loop [{ins[lo][0]:#x}..{ins[hi][0]:#x}] {kind}: {n} instr  ALU={alu} FMA={fma} other={other}")
        if kind == "inner":
            print("     " + "  ".join(f"{k}:{v}" for k, v in hist.most_common()))
