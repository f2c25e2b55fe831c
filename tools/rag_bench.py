#!/usr/bin/env python
"""Ragged vs uniform read batches on the packed 16-bit kernels (device-resident): python tools/rag_bench.py [pairs] [cfg ...]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gonomics_b200 import align  # noqa: E402
from gonomics_b200.synth import synth_pairs  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
cfgs = sys.argv[2:] or [""]
dev = torch.device("cuda:0")
a, sao, b, sbo = synth_pairs(20260111, P, 500, 150)
rr = np.random.default_rng(3)
S = align.HumanChimpTwoScoreMatrix
for name, (nlo, mlo) in (("uniform 500x150", (500, 150)), ("ragged 300-500 x 100-150", (300, 100)), ("ragged 480-500 x 141-150", (480, 141))):
    nl, ml = rr.integers(nlo, 501, size=P), rr.integers(mlo, 151, size=P)
    rao, rbo = np.zeros(P + 1, np.int64), np.zeros(P + 1, np.int64)
    np.cumsum(nl, out=rao[1:])
    np.cumsum(ml, out=rbo[1:])
    ia = np.repeat(sao[:-1] - rao[:-1], nl) + np.arange(rao[-1])
    ib = np.repeat(sbo[:-1] - rbo[:-1], ml) + np.arange(rbo[-1])
    ta, tb = torch.from_numpy(a[ia]).to(dev), torch.from_numpy(b[ib]).to(dev)
    tao, tbo = torch.from_numpy(rao).to(dev), torch.from_numpy(rbo).to(dev)
    sc = torch.zeros(P, dtype=torch.int64, device=dev)
    off = torch.zeros(P + 1, dtype=torch.int64, device=dev)
    cg = torch.zeros(P * 12 * 16, dtype=torch.uint8, device=dev)
    st = torch.zeros(1, dtype=torch.int32, device=dev)
    cells = int((nl * ml).sum())
    for cfg in cfgs:
        ctx = align.Context(0)
        for kv in cfg.split(","):
            if kv:
                k_, v_ = kv.split("=")
                ctx.set_option(k_, int(v_))
        for want in (False, True):
            best = 1e9
            for it in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ctx.batch_device(1, ta.data_ptr(), tao.data_ptr(), tb.data_ptr(), tbo.data_ptr(), rao, rbo, P, S, -600, -150, want,
                                 sc.data_ptr(), cg.data_ptr(), off.data_ptr(), P * 12, st.data_ptr(), torch.cuda.current_stream().cuda_stream)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            print(f"{name:28s} {cfg:24s} {'trace' if want else 'score'}: {best:8.2f} ms = {cells / best / 1e6:8.1f} GCUPS = {P / best / 1e3:6.1f} Mpairs/s "
                  f"path {ctx.last_kernel_path()}", flush=True)
        ctx.close()
