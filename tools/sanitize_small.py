#!/usr/bin/env python
"""Small instances of every round-2 kernel path, checked against the oracle; meant to run under
`compute-sanitizer --tool memcheck python tools/sanitize_small.py` (minutes, not hours)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as orc  # noqa: E402
from gonomics_b200 import align  # noqa: E402
from gonomics_b200.synth import pack_uniform, synth_pairs  # noqa: E402

S = orc.HUMAN_CHIMP_TWO_SCORE_MATRIX


def same(got, want, what):
    ok = np.array_equal(got[0], want[0]) and (got[1] is None or (np.array_equal(got[1], want[1]) and
                                                                np.array_equal(got[2]["op"], want[2]["op"]) and
                                                                np.array_equal(got[2]["run_length"], want[2]["run_length"])))
    print(("ok   " if ok else "FAIL ") + what, flush=True)
    return ok


ok = True
with align.Context(0) as c:
    c.set_option("chunk_pairs", 400)
    # uniform read-sized batch: bytes (small: no packing), 2-bit (TMA kernels, checkpoint path reading words), tail quad
    P, n, m = 1001, 300, 141
    a, ao, b, bo = synth_pairs(7, P, n, m)
    want = orc.batch(a, ao, b, bo, S, -600, -150, 1, True, 4)
    ok &= same(c.affine_gap_batch(a, ao, b, bo, S, -600, -150, True, True), want, "C3 bytes (fill16 CKPT + trace)")
    ok &= same(c.affine_gap_batch(a, ao, b, bo, S, -600, -150, True, False), (want[0], None, None), "C2 bytes (fill16, skew 2, cursor)")
    wa, wb = pack_uniform(a, P, n), pack_uniform(b, P, m)
    ok &= same(c.affine_gap_batch_twobit(wa, n, wb, m, S, -600, -150, True, True, n_pairs=P), want, "C3 2-bit (TMA, joint table)")
    ok &= same(c.affine_gap_batch_twobit(wa, n, wb, m, S, -600, -150, True, False, n_pairs=P), (want[0], None, None), "C2 2-bit")
    # packed while staged (>= 4096 pairs)
    P2 = 4100
    a2, ao2, b2, bo2 = synth_pairs(8, P2, 96, 64)
    c.set_option("chunk_pairs", 2000)
    ok &= same(c.affine_gap_batch(a2, ao2, b2, bo2, S, -600, -150, True, True), orc.batch(a2, ao2, b2, bo2, S, -600, -150, 1, True, 4),
               "byte batch packed while staged")
    # ragged batch on the packed kernels
    rng = np.random.default_rng(3)
    al = [rng.integers(0, 4, size=int(rng.integers(300, 360)), dtype=np.uint8) for _ in range(600)]
    be = [rng.integers(0, 4, size=int(rng.integers(100, 150)), dtype=np.uint8) for _ in range(600)]
    ac, aof = np.concatenate(al), np.concatenate([[0], np.cumsum([len(x) for x in al])]).astype(np.int64)
    bc, bof = np.concatenate(be), np.concatenate([[0], np.cumsum([len(x) for x in be])]).astype(np.int64)
    ok &= same(c.affine_gap_batch(ac, aof, bc, bof, S, -600, -150, True, True), orc.batch(ac, aof, bc, bof, S, -600, -150, 1, True, 4),
               "ragged (binned quads, CKPT)")
    # long pairs on the tile-checkpoint kernel
    c.set_option("long_ckpt", 1)
    al = [rng.integers(0, 4, size=2500, dtype=np.uint8) for _ in range(6)]
    be = [np.concatenate([x[:1200], x[1230:]]) for x in al]
    ac, aof = np.concatenate(al), np.arange(7, dtype=np.int64) * 2500
    bc, bof = np.concatenate(be), np.arange(7, dtype=np.int64) * 2470
    ok &= same(c.affine_gap_batch(ac, aof, bc, bof, S, -600, -150, False, True), orc.batch(ac, aof, bc, bof, S, -600, -150, 0, True, 4),
               "long pairs (affine_long_kernel)")
print("ALL OK" if ok else "FAILURES")
sys.exit(0 if ok else 1)
