#!/usr/bin/env python
"""Small driver for profiling seed_kernel: 262,144 reads x 150 bp against a 16 Mb reference."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gonomics_b200 import align, genomegraph  # noqa: E402

rng = np.random.default_rng(3)
g = rng.integers(0, 4, size=1 << 24, dtype=np.uint8)
ctx = align.Context(0)
ix = genomegraph.SeedIndex([g], 32, 32, ctx)
n = 262144
st = rng.integers(0, len(g) - 150, size=n)
reads = g[st[:, None] + np.arange(150)[None, :]]
mut = rng.random(reads.shape) < 0.02
reads[mut] = (reads[mut] + 1) % 4
seeds, off = ix.seed_batch(np.ascontiguousarray(reads.reshape(-1)), np.arange(n + 1, dtype=np.int64) * 150)
print("seeds per read", off[-1] / n)
