#!/usr/bin/env python
"""Kernel A/B timing on the GPU box: device-resident batches, fill-kernel time from the library's events.

    python tools/kbench.py [--pairs N] name=value[,name=value...] ...
Each positional argument is one configuration (gnx_set_option pairs)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gonomics_b200 import align  # noqa: E402
from gonomics_b200.synth import synth_pairs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=500_000)
ap.add_argument("--n", type=int, default=500)
ap.add_argument("--m", type=int, default=150)
ap.add_argument("--kind", type=int, default=1)
ap.add_argument("--check", action="store_true", help="diff the first --check-pairs pairs against the oracle")
ap.add_argument("--check-pairs", type=int, default=5000)
ap.add_argument("--workspace-gb", type=float, default=0)
ap.add_argument("--cap-per-pair", type=int, default=16)
ap.add_argument("--score-only", action="store_true")
ap.add_argument("--twobit", action="store_true", help="inputs as dnaTwoBit words (gnx_batch_device_twobit)")
ap.add_argument("--iters", type=int, default=4)
ap.add_argument("configs", nargs="*", default=["fill_impl=1", "fill_impl=2"])
args = ap.parse_args()
P = args.pairs
a, ao, b, bo = synth_pairs(20260102, P, args.n, args.m)
dev = torch.device("cuda:0")
ta, tb = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
tao, tbo = torch.from_numpy(ao).to(dev), torch.from_numpy(bo).to(dev)
if args.twobit:
    from gonomics_b200.synth import pack_uniform
    pad = torch.zeros(64, dtype=torch.int64, device=dev)
    twa = torch.cat([torch.from_numpy(pack_uniform(a, P, args.n).view(np.int64)).to(dev), pad])
    twb = torch.cat([torch.from_numpy(pack_uniform(b, P, args.m).view(np.int64)).to(dev), pad])
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
    k = min(P, args.check_pairs)
    ref = orc.batch(a[:k * args.n], ao[:k + 1], b[:k * args.m], bo[:k + 1], orc.HUMAN_CHIMP_TWO_SCORE_MATRIX, -600, -150,
                    args.kind, True, os.cpu_count())
for cfg in args.configs:
    ctx = align.Context(0, int(args.workspace_gb * (1 << 30)))
    for kv in cfg.split(","):
        if kv:
            k_, v_ = kv.split("=")
            ctx.set_option(k_, int(v_))
    for want in ((False,) if args.score_only else (False, True)):
        best_fill, best_tot = 1e9, 1e9
        for it in range(args.iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if args.twobit:
                ctx.batch_device_twobit(args.kind, twa.data_ptr(), args.n, twb.data_ptr(), args.m, P, S, -600, -150, want,
                                        score.data_ptr(), cig.data_ptr(), off.data_ptr(), cap, status.data_ptr(),
                                        torch.cuda.current_stream().cuda_stream)
            else:
                ctx.batch_device(args.kind, ta.data_ptr(), tao.data_ptr(), tb.data_ptr(), tbo.data_ptr(), ao, bo, P, S, -600, -150,
                                 want, score.data_ptr(), cig.data_ptr(), off.data_ptr(), cap, status.data_ptr(),
                                 torch.cuda.current_stream().cuda_stream)
            e1.record()
            torch.cuda.synchronize()
            fill_ms, _, _ = ctx.last_fill_stats()
            best_fill, best_tot = min(best_fill, fill_ms), min(best_tot, e0.elapsed_time(e1))
        ok = ""
        if ref is not None:
            k = len(ref[0])
            good = np.array_equal(score[:k].cpu().numpy(), ref[0])
            if want:
                good = good and np.array_equal(off[:k + 1].cpu().numpy(), ref[1])
                got = cig.cpu().numpy()[:int(ref[1][-1]) * 16].view(align.CIGAR_DTYPE)
                good = good and np.array_equal(got["run_length"], ref[2]["run_length"]) and np.array_equal(got["op"], ref[2]["op"])
            ok = "  parity=" + ("OK" if good else "FAIL")
        print(f"{cfg:40s} {'trace' if want else 'score'}: fill {best_fill:8.2f} ms = {cells / best_fill / 1e6:8.1f} GCUPS | "
              f"total {best_tot:8.2f} ms = {cells / best_tot / 1e6:8.1f} GCUPS{ok}", flush=True)
    ctx.close()
