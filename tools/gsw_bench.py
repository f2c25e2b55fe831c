#!/usr/bin/env python
"""gsw driver timing on the GPU box: python tools/gsw_bench.py [pairs] [genome_bases]  (GNX_GSW_TIMING=1 prints the phases)"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gonomics_b200 import align, genomegraph  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 19
g_len = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 26
rng = np.random.default_rng(5)
genome = rng.integers(0, 4, size=g_len, dtype=np.uint8)
ctx = align.Context(0)
for kv in os.environ.get("GSW_OPTS", "").split(","):  # e.g. GSW_OPTS=chunk_pairs=1048576
    if "=" in kv:
        ctx.set_option(kv.split("=")[0], int(kv.split("=")[1]))
ix = genomegraph.SeedIndex([genome], 32, 32, ctx)
start = rng.integers(0, g_len - 600, size=pairs)
frag = rng.integers(300, 500, size=pairs)
ar = np.arange(150)
reads = np.empty((2 * pairs, 150), dtype=np.uint8)
reads[0::2] = genome[start[:, None] + ar[None, :]]
reads[1::2] = (3 - genome[(start + frag - 150)[:, None] + ar[None, :]])[:, ::-1]
mut = rng.random(reads.shape) < 0.01
reads[mut] = (reads[mut] + rng.integers(1, 4, size=int(mut.sum()), dtype=np.uint8)) % 4
cat = np.ascontiguousarray(reads.reshape(-1))
off = np.arange(2 * pairs + 1, dtype=np.int64) * 150
S = align.HumanChimpTwoScoreMatrix
genomegraph.gsw_batch(ix, cat[:150 * 2000], off[:2001], S, paired=True)
for it in range(3):
    t0 = time.perf_counter()
    recs, cig = genomegraph.gsw_batch(ix, cat, off, S, paired=True, cigar_cap=8 * pairs)
    dt = time.perf_counter() - t0
    print(f"{pairs} pairs: {dt * 1e3:.1f} ms = {pairs / dt / 1e6:.2f} Mpairs/s; mapped {(recs['aln_score'] >= 1200).mean():.3f}", flush=True)
