#!/usr/bin/env python
"""One launch of twobit_pack_kernel over 2^31 bases (for the ncu capture behind profiles/ncu_traffic.json)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gonomics_b200 import align  # noqa: E402

ctx = align.Context(0)
n = 1 << 31
seq = torch.randint(0, 4, (n,), dtype=torch.uint8, device="cuda:0")
words = torch.zeros(n // 32, dtype=torch.int64, device="cuda:0")
for _ in range(2):
    ctx._check(ctx._L.gnx_twobit_pack_device(ctx._h, seq.data_ptr(), n, 0, words.data_ptr(), torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
