#!/usr/bin/env python
"""Long-pair (C4-shaped) A/B timing: device-resident batch of n x m global pairs with CIGAR; each configuration is a
comma-separated list of gnx_set_option pairs.  python tools/kbench_long.py [--pairs N] [--n 10000] [--m 10000] cfg ..."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gonomics_b200 import align  # noqa: E402
from gonomics_b200.synth import synth_pairs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=2368)
ap.add_argument("--n", type=int, default=10000)
ap.add_argument("--m", type=int, default=10000)
ap.add_argument("--kind", type=int, default=0)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--workspace-gb", type=float, default=0)
ap.add_argument("--cap-per-pair", type=int, default=600)
ap.add_argument("--check", type=int, default=0, help="diff this many pairs against the oracle")
ap.add_argument("configs", nargs="*", default=["long_ckpt=1"])
args = ap.parse_args()
P = args.pairs
a, ao, b, bo = synth_pairs(20260104, P, args.n, args.m)
dev = torch.device("cuda:0")
ta, tb = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
tao, tbo = torch.from_numpy(ao).to(dev), torch.from_numpy(bo).to(dev)
score = torch.zeros(P, dtype=torch.int64, device=dev)
off = torch.zeros(P + 1, dtype=torch.int64, device=dev)
cap = P * args.cap_per_pair
cig = torch.zeros(cap * 16, dtype=torch.uint8, device=dev)
status = torch.zeros(1, dtype=torch.int32, device=dev)
S = align.HumanChimpTwoScoreMatrix
cells = P * args.n * args.m
ref = None
if args.check:
    import oracle as orc
    k = min(P, args.check)
    ref = orc.batch(a[:ao[k]], ao[:k + 1], b[:bo[k]], bo[:k + 1], orc.HUMAN_CHIMP_TWO_SCORE_MATRIX, -600, -150, args.kind, True,
                    os.cpu_count())
for cfg in args.configs:
    ctx = align.Context(0, int(args.workspace_gb * (1 << 30)))
    for kv in cfg.split(","):
        if kv:
            k_, v_ = kv.split("=")
            ctx.set_option(k_, int(v_))
    best_fill, best_tot = 1e9, 1e9
    for it in range(args.reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ctx.batch_device(args.kind, ta.data_ptr(), tao.data_ptr(), tb.data_ptr(), tbo.data_ptr(), ao, bo, P, S, -600, -150, True,
                         score.data_ptr(), cig.data_ptr(), off.data_ptr(), cap, status.data_ptr(),
                         torch.cuda.current_stream().cuda_stream)
        e1.record()
        torch.cuda.synchronize()
        fill_ms, _, _ = ctx.last_fill_stats()
        if it > 0:
            best_fill, best_tot = min(best_fill, fill_ms), min(best_tot, e0.elapsed_time(e1))
    ok = ""
    if ref is not None:
        k = len(ref[0])
        good = np.array_equal(score[:k].cpu().numpy(), ref[0]) and np.array_equal(off[:k + 1].cpu().numpy(), ref[1])
        got = cig.cpu().numpy()[:int(ref[1][-1]) * 16].view(align.CIGAR_DTYPE)
        good = good and np.array_equal(got["run_length"], ref[2]["run_length"]) and np.array_equal(got["op"], ref[2]["op"])
        ok = "  parity=" + ("OK" if good else "FAIL")
    free_b, tot_b = torch.cuda.mem_get_info(0)
    print(f"{cfg:40s} {P} pairs {args.n}x{args.m}: fill {best_fill:8.2f} ms = {cells / best_fill / 1e6:8.1f} GCUPS | total "
          f"{best_tot:8.2f} ms = {cells / best_tot / 1e6:8.1f} GCUPS | status {int(status.item())} "
          f"cigar elems {int(off[-1].item())} | device mem in use {(tot_b - free_b) / 1e9:.1f} GB{ok}", flush=True)
    ctx.close()
