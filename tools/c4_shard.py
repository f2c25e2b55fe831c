#!/usr/bin/env python
"""BASELINE config C4 at per-GPU scale: 100k pairs of 10 kb x 10 kb over 8 GPUs = 12,500 pairs per GPU, global affine
with full CIGAR, through the host-buffer API (H2D + D2H inside the timed region).  One GPU's shard is run here."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gonomics_b200 import align  # noqa: E402
from gonomics_b200.synth import synth_pairs  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 12_500
N = M = 10_000
a, ao, b, bo = synth_pairs(20260104, P, N, M)
free_b, _ = torch.cuda.mem_get_info(0)
ctx = align.Context(0, int(free_b * 0.8))
S = align.HumanChimpTwoScoreMatrix
t0 = time.perf_counter()
sc, off, cig = ctx.affine_gap_batch(a, ao, b, bo, S, -600, -150, False, True, cigar_cap=P * 600)
dt = time.perf_counter() - t0
cells = P * N * M
print(f"C4 shard: {P} pairs 10kb x 10kb global + CIGAR, host API: {dt:.2f} s = {cells / dt / 1e9:.1f} GCUPS; "
      f"{int(off[-1])} cigar elements ({off[-1] / P:.1f} per pair), mean score {sc.mean():.0f}, launches {ctx.launch_count}")
# size-independent properties: every cigar consumes exactly n and m bases and runs are maximal
rl, op = cig["run_length"], cig["op"]
cons_a = np.add.reduceat(np.where(op != 1, rl, 0), off[:-1])
cons_b = np.add.reduceat(np.where(op != 2, rl, 0), off[:-1])
assert np.all(cons_a == N) and np.all(cons_b == M), "a cigar does not consume the whole pair"
same = op[1:] == op[:-1]
same[off[1:-1] - 1] = False
assert not same.any(), "adjacent runs with the same op"
if P >= 4:
    import oracle as orc
    for p in (0, P // 2, P - 1):
        osc, ocig = orc.affine_gap_highmem(a[ao[p]:ao[p + 1]], b[bo[p]:bo[p + 1]], orc.HUMAN_CHIMP_TWO_SCORE_MATRIX, -600, -150)
        got = [(int(r), int(o)) for r, o in cig[off[p]:off[p + 1]]]
        assert osc == int(sc[p]) and ocig == got, f"pair {p} differs from the oracle"
    print("parity: 3 pairs diffed against the oracle (score + cigar): OK; all cigars consume n and m, runs maximal")
ctx.close()
