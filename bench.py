#!/usr/bin/env python
"""bench.py -- GCUPS of the affine-gap DP hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--pairs P] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic pairs resident in HBM.
Headline workload (BASELINE.json configs[2], "C3"): 10^7 pairs IN TOTAL, target 500 x query 150, AffineGapLocal
semantics (free end gaps), HumanChimpTwo matrix, O=-600, E=-150, full traceback + CIGAR, batch-sharded over the N
GPUs (strong scaling; `--scaling weak` keeps 10^7 pairs per GPU), inputs as dnaTwoBit words, and -- for N > 1 --
the NCCL gather of every shard's scores AND cigars inside the timed step.  The same line carries configs[1]
("C2": score only) under "score_only", configs[3] ("C4": 12,500 pairs of 10 kb x 10 kb per GPU) and the other
shapes under "other_workloads".  GCUPS counts each DP cell once: sum(n*m) / seconds / 1e9.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_LEN, M_LEN = 500, 150
GAP_OPEN, GAP_EXTEND = -600, -150
SEED = 20260102
# algorithmic bytes: inputs 2-bit packed + 16 B offsets + 8 B score per pair;
# traceback adds 0.75 B per cell (three 2-bit source-plane codes)
BYTES_PER_PAIR_SCORE = (N_LEN + 3) // 4 + (M_LEN + 3) // 4 + 16 + 8
BYTES_PER_CELL_TRACE = 0.75
WORKLOAD = ("C3 (BASELINE.json configs[2]): 10^7 pairs, target 500 x query 150, semi-global affine gap (AffineGapLocal), "
            "full traceback + CIGAR, batch-sharded over the GPUs")


def ncu_traffic(kernel: str, pairs_per_launch: float):
    """DRAM bytes per launch measured by ncu (profiles/ncu_traffic.json), scaled to this run's launch size."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)
        return t[kernel]["dram_bytes"] * pairs_per_launch / t[kernel].get("units_per_launch", t["pairs_per_launch"])
    except Exception:
        return None


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md's clocks line).  NVML is polled from a
    thread every 5 ms (an `nvidia-smi -lms` child needs ~0.3 s before its first row, longer than a multi-GPU timed
    region); `nvidia-smi` is the fallback when the NVML binding is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    # NVML clocks-event-reason bits (nvml.h: nvmlClocksEventReason*)
    BITS = (("sw_power_cap", 0x4), ("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40))

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.nv = index, [], None, None
        self.sm, self.bits, self.mx, self.stop_flag = [], 0, None, False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.nv = (pynvml, h)
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv, h = self.nv
        reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.bits |= int(reasons(h))
            except Exception:
                pass
            time.sleep(0.005)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nv:
            self.stop_flag = True
            self.t.join(timeout=2)
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.mx,
                    "reasons": sorted(n for n, b in self.BITS if self.bits & b), "samples": len(self.sm), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def cpu_reference_gcups(n_pairs: int, want_cigar: bool, threads: int, first_pair: int = 0):
    """The oracle (C restatement of the Go path) on the host cores: the reported CPU baseline."""
    import oracle as orc
    from gonomics_b200.synth import synth_pairs
    a, ao, b, bo = synth_pairs(SEED, n_pairs, N_LEN, M_LEN, first_pair=first_pair)
    t0 = time.perf_counter()
    orc.batch(a, ao, b, bo, orc.HUMAN_CHIMP_TWO_SCORE_MATRIX, GAP_OPEN, GAP_EXTEND, 1, want_cigar, threads)
    dt = time.perf_counter() - t0
    return n_pairs * N_LEN * M_LEN / dt / 1e9, dt


def twobit_block(ctx, L, dev, stream, rank, world, barrier, args):
    """dnaTwoBit.NewTwoBit of one long sequence (device-resident, HBM roofline) and genomeGraph.seedMapMemPool
    over a synthetic linear reference (host-buffer API), with the oracle timed beside it on rank 0."""
    import torch
    import torch.distributed as dist
    from gonomics_b200 import genomegraph
    out = {}
    peak, peak_src = hbm_peak()
    # -- pack: 2^31 bases in HBM -> 2^26 words; algorithmic bytes = 1 B/base in + 0.25 B/base out
    n = 1 << 31
    seq = torch.randint(0, 4, (n,), dtype=torch.uint8, device=dev)
    words = torch.zeros(n // 32, dtype=torch.int64, device=dev)
    for _ in range(3):
        ctx._check(L.gnx_twobit_pack_device(ctx._h, seq.data_ptr(), n, 0, words.data_ptr(), stream))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        ctx._check(L.gnx_twobit_pack_device(ctx._h, seq.data_ptr(), n, 0, words.data_ptr(), stream))
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / reps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    gbs = 1.25 * n / (ms * 1e-3) / 1e9
    # size-independent check at full size: base i of the packing is seq[i] (sampled) and a word checksum
    idx = torch.randint(0, n, (1 << 16,), device=dev)
    got = (words[idx // 32] >> (62 - 2 * (idx % 32))) & 3
    assert torch.equal(got.to(torch.uint8), seq[idx]), "2-bit packing differs from the input bases"
    out["twobit_pack_2Gbase"] = {
        "value": world * n / (ms * 1e-3) / 1e9, "unit": "Gbases/s", "ms_per_step": ms,
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                     "traffic": ncu_traffic("twobit_pack_kernel", n), "peak_source": peak_src, "kernel": "twobit_pack_kernel",
                     "algorithmic_bytes_per_launch": 1.25 * n},
        "note": "dnaTwoBit.NewTwoBit of one 2^31-base sequence resident in HBM (inputs exceed L2)"}
    del seq, words, idx, got
    torch.cuda.empty_cache()
    # -- seeds: 1M reads x 150 bp against a 64 Mb linear reference, seedLen 32 / seedStep 32 (cmd/gsw defaults)
    rng = np.random.default_rng(SEED + 11 + rank)
    g_len, n_reads, r_len = 1 << 26, 1_000_000, 150
    genome = rng.integers(0, 4, size=g_len, dtype=np.uint8)
    t0 = time.perf_counter()
    ix = genomegraph.SeedIndex([genome], 32, 32, ctx)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    starts = rng.integers(0, g_len - r_len, size=n_reads)
    reads = genome[starts[:, None] + np.arange(r_len)[None, :]]
    mut = rng.random(reads.shape) < 0.02
    reads[mut] = (reads[mut] + rng.integers(1, 4, size=int(mut.sum()), dtype=np.uint8)) % 4
    flip = rng.random(n_reads) < 0.5
    reads[flip] = (3 - reads[flip])[:, ::-1]
    import ctypes as C
    from gonomics_b200._lib import SEED_DTYPE

    def pinned(count, dtype):
        dt_ = np.dtype(dtype)
        ptr = L.gnx_host_alloc(count * dt_.itemsize)
        raw = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(count * dt_.itemsize,))
        return ptr, raw.view(dt_)
    p1, rcat = pinned(n_reads * r_len, np.uint8)  # reads and results in page-locked memory: direct DMA
    rcat[:] = reads.reshape(-1)
    roff = np.arange(n_reads + 1, dtype=np.int64) * r_len
    p2, seed_buf = pinned(8 * n_reads, SEED_DTYPE)
    p3, soff_buf = pinned(n_reads + 1, np.int64)
    seeds, soff = ix.seed_batch(rcat, roff, out=(seed_buf, soff_buf))
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        seeds, soff = ix.seed_batch(rcat, roff, out=(seed_buf, soff_buf))
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    blk = {"value": world * n_reads / dt / 1e6, "unit": "Mreads/s", "ms_per_step": dt * 1e3,
           "seeds_per_read": float(soff[-1]) / n_reads, "index_entries": ix.n_entries, "index_build_s": build_s,
           "note": "genomeGraph.seedMapMemPool, host-buffer API (pinned buffers; H2D reads + D2H seeds inside the timed region), "
                   "1M reads x 150 bp, 2 % substitutions, both strands, 64 Mb reference, seedLen 32 / step 32"}
    if rank == 0 and world == 1 and not args.no_cpu:
        import oracle as orc
        key, loc = ix.entries()
        off = np.array([0, g_len], dtype=np.int64)
        k = 20000
        packed = orc.pack_nodes(genome, off)  # Node.SeqTwoBit, built once like the reference's genome graph
        t0 = time.perf_counter()
        ok = True
        for r in range(k):
            want = orc.seeds_for_read(key, loc, genome, off, reads[r], 32, packed)
            got = seeds[soff[r]:soff[r + 1]]
            ok &= len(got) == len(want) and all(np.array_equal(got[f], want[:, c]) for c, f in enumerate(got.dtype.names))
        blk["cpu_baseline"] = {"value": k / (time.perf_counter() - t0) / 1e6, "unit": "Mreads/s", "cores": 1, "kind": "port",
                               "sample": f"first {k} reads, one thread (C restatement of seedMapMemPool incl. the read's rainbow "
                                         "tables; ctypes call per read)"}
        blk["parity_spot_check"] = bool(ok)
    out["gsw_seeds_1M_reads_150bp"] = blk
    seeds = soff = seed_buf = soff_buf = rcat = None
    for ptr in (p1, p2, p3):
        L.gnx_host_free(ptr)
    # -- the gsw extend step on read-sized flanks: LeftDynamicAln / RightDynamicAln, 1M pairs each, target window
    #    175 (extension = perfectScore/600 + len(read) - seed, genomeGraph/toGiraf.go:32) x read flank 100
    from gonomics_b200 import align as _al
    from gonomics_b200._lib import GNX_EXT_LEFT, GNX_EXT_RIGHT
    n_ext, tn, qm = 1_000_000, 175, 100
    est = rng.integers(0, g_len - tn - 8, size=n_ext)
    tgt = genome[est[:, None] + np.arange(tn)[None, :]]
    qry = tgt[:, tn - qm:].copy()  # left side: the flank ends where the window ends
    mutq = rng.random(qry.shape) < 0.03
    qry[mutq] = (qry[mutq] + rng.integers(1, 4, size=int(mutq.sum()), dtype=np.uint8)) % 4
    from gonomics_b200._lib import CIGAR_DTYPE
    ptrs = []

    def pin(count, dtype):
        ptr, arr = pinned(count, dtype)
        ptrs.append(ptr)
        return arr
    tcat, toff = pin(n_ext * tn, np.uint8), np.arange(n_ext + 1, dtype=np.int64) * tn
    tcat[:] = tgt.reshape(-1)
    qcat, qoff = pin(n_ext * qm, np.uint8), np.arange(n_ext + 1, dtype=np.int64) * qm
    qcat[:] = qry.reshape(-1)
    ext_out = (pin(n_ext, np.int64), pin(n_ext, np.int64), pin(n_ext, np.int64), pin(n_ext + 1, np.int64),
               pin(8 * n_ext, CIGAR_DTYPE))
    ext = {}
    for name, side in (("left", GNX_EXT_LEFT), ("right", GNX_EXT_RIGHT)):
        if side == GNX_EXT_RIGHT:  # right side: the flank starts where the window starts
            qry = tgt[:, :qm].copy()
            qry[mutq] = (qry[mutq] + 1) % 4
            qcat[:] = qry.reshape(-1)
        ctx.extend_batch(side, tcat, toff, qcat, qoff, _al.HumanChimpTwoScoreMatrix, -600, out=ext_out)
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            sc, ei, ej, coff, cig = ctx.extend_batch(side, tcat, toff, qcat, qoff, _al.HumanChimpTwoScoreMatrix, -600,
                                                     out=ext_out)
        torch.cuda.synchronize()
        dte = (time.perf_counter() - t0) / 3
        if world > 1:
            t = torch.tensor([dte], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dte = float(t.item())
        ext[name] = {"Mpairs_per_s": world * n_ext / dte / 1e6, "GCUPS": world * n_ext * tn * qm / dte / 1e9,
                     "ms_per_step": dte * 1e3, "mean_score": float(sc.mean())}
        if rank == 0 and world == 1 and not args.no_cpu:
            import oracle as orc
            fn = orc.left_dynamic_aln if side == GNX_EXT_LEFT else orc.right_dynamic_aln
            ok = True
            for r in range(300):
                w = fn(tgt[r], qry[r], orc.HUMAN_CHIMP_TWO_SCORE_MATRIX, -600)
                g = [(int(x["run_length"]), chr(int(x["op"]))) for x in cig[coff[r]:coff[r + 1]]]
                ok &= (w[0] == int(sc[r]) and w[1] == g and w[2] == int(ei[r]) and w[3] == int(ej[r]))
            ext[name]["parity_spot_check"] = bool(ok)
    out["gsw_extend_1M_pairs_175x100"] = {
        "value": ext["left"]["Mpairs_per_s"], "unit": "Mpairs/s (LeftDynamicAln)", "left": ext["left"], "right": ext["right"],
        "note": "genomeGraph.LeftDynamicAln / RightDynamicAln with route, host-buffer API (pinned buffers, "
                "H2D + D2H inside the timed region), linear gap -600, HumanChimpTwo"}
    sc = ei = ej = coff = cig = ext_out = tcat = qcat = None
    for ptr in ptrs:
        L.gnx_host_free(ptr)
    ix.close()
    return out



def gsw_block(ctx, L, dev, rank, world, args):
    """BASELINE configs[4] shape ("C5"): paired-end 2 x 150 bp reads through the whole gsw per-read driver
    (gnx_gsw_batch: seeds -> candidate ordering -> left/right extension DPs -> replay of GraphSmithWatermanToGiraf /
    WrapPairGiraf) against a synthetic >= 1 Gb linear reference, host buffers in and out, blocks of 2^20 reads."""
    import torch
    import torch.distributed as dist
    from gonomics_b200 import align as _al
    from gonomics_b200 import genomegraph
    rng = np.random.default_rng(SEED + 55)  # the same genome on every rank
    g_len = args.gsw_genome
    t0 = time.perf_counter()
    genome = rng.integers(0, 4, size=g_len, dtype=np.uint8)
    ix = genomegraph.SeedIndex([genome], 32, 32, ctx)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    pairs_total = args.gsw_pairs
    lo, hi = pairs_total * rank // world, pairs_total * (rank + 1) // world
    n_pairs = hi - lo
    r_len, blk_pairs = 150, 1 << 19
    rng = np.random.default_rng(SEED + 56 + rank)

    def make_block(np_):
        """np_ fragments of 300-500 bases; mates from the two ends on opposite strands; 1 % substitutions, a short
        indel in 10 % of the reads; 2 % of the reads unrelated."""
        start = rng.integers(0, g_len - 600, size=np_)
        frag = rng.integers(300, 500, size=np_)
        ar = np.arange(r_len)
        fwd = genome[start[:, None] + ar[None, :]]
        rev = genome[(start + frag - r_len)[:, None] + ar[None, :]]
        rev = (3 - rev)[:, ::-1]
        reads = np.empty((2 * np_, r_len), dtype=np.uint8)
        reads[0::2], reads[1::2] = fwd, rev
        mut = rng.random(reads.shape) < 0.01
        reads[mut] = (reads[mut] + rng.integers(1, 4, size=int(mut.sum()), dtype=np.uint8)) % 4
        ind = np.nonzero(rng.random(2 * np_) < 0.10)[0]
        pos = rng.integers(40, r_len - 40, size=len(ind))
        for r, p in zip(ind[::2], pos[::2]):      # deletion of one base (shift left, random base at the end)
            reads[r, p:-1] = reads[r, p + 1:]
        for r, p in zip(ind[1::2], pos[1::2]):    # insertion of one base
            reads[r, p + 1:] = reads[r, p:-1].copy()
            reads[r, p] = (reads[r, p] + 1) % 4
        junk = np.nonzero(rng.random(2 * np_) < 0.02)[0]
        reads[junk] = rng.integers(0, 4, size=(len(junk), r_len), dtype=np.uint8)
        return np.ascontiguousarray(reads.reshape(-1))

    S = _al.HumanChimpTwoScoreMatrix
    blocks = []
    left = n_pairs
    while left > 0:
        b = min(left, blk_pairs)
        blocks.append((make_block(b), b))
        left -= b
    off_full = np.arange(2 * blk_pairs + 1, dtype=np.int64) * r_len
    genomegraph.gsw_batch(ix, blocks[0][0][:2 * 4096 * r_len], off_full[:2 * 4096 + 1], S, paired=True)  # warm-up
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mapped = proper = n_cig = 0
    first = None
    for cat, b in blocks:
        recs, cig = genomegraph.gsw_batch(ix, cat, off_full[:2 * b + 1], S, paired=True, cigar_cap=8 * b)
        mapped += int((recs["aln_score"] >= 1200).sum())
        proper += int((recs["flag"][1::2] & 1).sum())
        n_cig += len(cig)
        if first is None:
            first = (recs, cig)
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    out = {"value": pairs_total / dt / 1e6, "unit": "Mpairs/s (2 x 150 bp)", "seconds": dt, "pairs_total": pairs_total,
           "pairs_per_gpu": n_pairs, "reference_bases": g_len, "index_entries": ix.n_entries, "index_build_s": build_s,
           "mapped_fraction": mapped / max(2 * n_pairs, 1), "proper_pair_fraction": proper / max(n_pairs, 1),
           "h2d_bytes_per_step": int(2 * n_pairs * r_len), "d2h_bytes_per_step": int(2 * n_pairs * 48 + 16 * n_cig),
           "api": "gnx_gsw_batch (host buffers; seeds + extensions on the GPU, ordering / replay on the host threads)",
           "note": "BASELINE configs[4] shape on a synthetic linear reference (the config's 3 Gb / 100 M pairs scaled to "
                   "what one default bench run holds); reads sharded over the ranks"}
    if rank == 0 and not args.no_cpu:
        from oracle import gsw as ogsw
        k = 400
        if True:
            gg = ogsw.LinearGenome([genome], 32, 32)  # the oracle's own seed map of the whole reference (~13 s per Gb)
            recs, cig = first
            cat = blocks[0][0]
            t0 = time.perf_counter()
            ok = True
            for p in range(k // 2):
                wf, wr = ogsw.wrap_pair_giraf(gg, cat[(2 * p) * r_len:(2 * p + 1) * r_len], cat[(2 * p + 1) * r_len:(2 * p + 2) * r_len], S)
                for r, w in ((2 * p, wf), (2 * p + 1, wr)):
                    g = recs[r]
                    ok &= (int(g["aln_score"]), int(g["t_start"]), int(g["t_end"]), int(g["flag"]), bool(g["pos_strand"])) == \
                          (w.AlnScore, w.TStart, w.TEnd, w.Flag, w.PosStrand)
            out["cpu_baseline"] = {"value": (k // 2) / (time.perf_counter() - t0) / 1e6, "unit": "Mpairs/s", "cores": 1, "kind": "port",
                                   "sample": f"first {k // 2} pairs, one thread (oracle/gsw.py: the sequential restatement of the "
                                             "reference loop over the C restatements of seedMapMemPool / Left/RightDynamicAln)"}
            out["parity_spot_check"] = bool(ok)
    ix.close()
    return out


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The Go toolchain is not in
    this image, so this is the C restatement (oracle/), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = max(2000, 4000 * threads)  # ~3 s of host work per step
    vals = []
    for i in range(args.warmup + args.steps):
        g, dt = cpu_reference_gcups(sample, True, threads, first_pair=i * sample)
        if i >= args.warmup:
            vals.append((g, dt))
    gcups = sample * N_LEN * M_LEN * len(vals) / sum(d for _, d in vals) / 1e9
    line = {
        "impl": "reference", "metric": "GCUPS", "value": gcups, "unit": "GCUPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(d for _, d in vals) / len(vals),
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "matrix": "HumanChimpTwo", "gap_open": GAP_OPEN, "gap_extend": GAP_EXTEND},
        "cpu_baseline": {"value": gcups, "unit": "GCUPS", "cores": threads, "kind": "port",
                         "sample": f"{sample} pairs of the C3 batch per step, traceback + cigar (C restatement of the Go "
                                   "path: affineGap_highMem incl. its per-call trace allocation; Go toolchain absent)"},
        "e2e": {"value": gcups, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_RESULT_FD = 1


def emit(line: dict):
    os.write(_RESULT_FD, (json.dumps(line) + "\n").encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pairs", type=int, default=10_000_000, help="pairs per step: in total (strong) or per GPU (weak)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--impl", default="gnx", choices=["gnx", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the C1 / C4 / const-gap / 2-bit / gsw blocks")
    ap.add_argument("--quick", action="store_true", help="profiling runs: headline only, warm-up not clamped")
    ap.add_argument("--c4-pairs", type=int, default=12_500, help="10 kb x 10 kb pairs per GPU in the C4 block")
    ap.add_argument("--gsw-pairs", type=int, default=10_000_000, help="read pairs (in total) of the gsw / C5 block")
    ap.add_argument("--gsw-genome", type=int, default=1 << 30, help="bases of the synthetic reference of the gsw block")
    args = ap.parse_args()
    # one rank per GPU shares the box's cores: cap the library's host threads (staging, packing, gsw host phases)
    lws = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))
    if lws > 1 and "GNX_HOST_THREADS" not in os.environ:
        os.environ["GNX_HOST_THREADS"] = str(max(2, (os.cpu_count() or 1) // lws))
    # stdout carries exactly ONE line, the JSON result: everything libraries print there (e.g. "NCCL version ..."
    # under torchrun) is sent to stderr by pointing fd 1 at fd 2 for the duration of the run.
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
        return
    if args.quick:
        args.no_e2e = args.no_cpu = args.no_extra = True
    else:
        args.warmup = max(args.warmup, 3)

    import ctypes as C

    import torch
    import torch.distributed as dist
    from gonomics_b200 import align, shard
    from gonomics_b200._lib import CIGAR_DTYPE, load
    from gonomics_b200.synth import pack_uniform, synth_pairs

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- this rank's shard: contiguous pair range [lo, hi) of the global batch ------------------------
    if args.scaling == "strong":
        lo, hi = args.pairs * rank // world, args.pairs * (rank + 1) // world
        total_pairs = args.pairs
    else:
        lo, hi = args.pairs * rank, args.pairs * (rank + 1)
        total_pairs = args.pairs * world
    P = hi - lo
    S = align.HumanChimpTwoScoreMatrix
    L = load()
    ctx = align.Context(local)
    cells = P * N_LEN * M_LEN            # this rank
    cells_all = total_pairs * N_LEN * M_LEN
    WN, WM = (N_LEN + 31) // 32, (M_LEN + 31) // 32

    pinned_ptrs = []

    def pinned_array(count, dtype):
        """Page-locked host array (gnx_host_alloc): DMA'd without a staging copy."""
        dt = np.dtype(dtype)
        ptr = L.gnx_host_alloc(max(count * dt.itemsize, 1))
        if not ptr:
            raise MemoryError("gnx_host_alloc failed")
        pinned_ptrs.append(ptr)
        raw = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(max(count * dt.itemsize, 1),))
        return raw[:count * dt.itemsize].view(dt)

    # ---- synthetic inputs (SURVEY 8d recipe) in pinned host memory: bytes and dnaTwoBit words -----------
    na, nb = P * N_LEN, P * M_LEN
    h_alpha, h_beta = pinned_array(na, np.uint8), pinned_array(nb, np.uint8)
    t0 = time.perf_counter()
    _, ao, _, bo = synth_pairs(SEED, P, N_LEN, M_LEN, first_pair=lo, alpha_out=h_alpha, beta_out=h_beta)
    h_wa, h_wb = pinned_array(P * WN, np.uint64), pinned_array(P * WM, np.uint64)
    pack_uniform(h_alpha, P, N_LEN, out=h_wa)
    pack_uniform(h_beta, P, M_LEN, out=h_wb)
    gen_s = time.perf_counter() - t0
    d_alpha, d_beta = torch.from_numpy(h_alpha).to(dev), torch.from_numpy(h_beta).to(dev)
    d_ao, d_bo = torch.from_numpy(ao).to(dev), torch.from_numpy(bo).to(dev)
    pad = torch.zeros(64, dtype=torch.int64, device=dev)  # the TMA of a tail quad reads a whole quad's words
    d_wa = torch.cat([torch.from_numpy(h_wa.view(np.int64)).to(dev), pad])
    d_wb = torch.cat([torch.from_numpy(h_wb.view(np.int64)).to(dev), pad])
    d_score = torch.zeros(P, dtype=torch.int64, device=dev)
    d_status = torch.zeros(1, dtype=torch.int32, device=dev)
    cig_cap = P * 12
    d_cig = torch.zeros(cig_cap * 16, dtype=torch.uint8, device=dev)
    d_off = torch.zeros(P + 1, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    gather_bytes = [0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step(want_cigar: bool, twobit: bool, gather: bool = True):
        if twobit:
            ctx.batch_device_twobit(1, d_wa.data_ptr(), N_LEN, d_wb.data_ptr(), M_LEN, P, S, GAP_OPEN, GAP_EXTEND, want_cigar,
                                    d_score.data_ptr(), d_cig.data_ptr() if want_cigar else 0,
                                    d_off.data_ptr() if want_cigar else 0, cig_cap, d_status.data_ptr(), stream)
        else:
            ctx.batch_device(1, d_alpha.data_ptr(), d_ao.data_ptr(), d_beta.data_ptr(), d_bo.data_ptr(), ao, bo, P, S,
                             GAP_OPEN, GAP_EXTEND, want_cigar, d_score.data_ptr(), d_cig.data_ptr() if want_cigar else 0,
                             d_off.data_ptr() if want_cigar else 0, cig_cap, d_status.data_ptr(), stream)
        if world > 1 and gather:  # the path's only exchange (north_star): every shard's scores -- and cigars -- to every rank
            res = shard.gather_device(d_score, d_off if want_cigar else None, d_cig if want_cigar else None, 16)
            gather_bytes[0] = res[4]

    def timed_device(want_cigar: bool, twobit: bool, steps=None, warmup=None):
        steps, warmup = steps or args.steps, args.warmup if warmup is None else warmup
        for _ in range(warmup):
            device_step(want_cigar, twobit)
        barrier()
        launches0 = ctx.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(local)
        sampler.start()
        e0.record()
        for _ in range(steps):
            device_step(want_cigar, twobit)
        e1.record()
        barrier()
        clocks = sampler.stop()
        ms = e0.elapsed_time(e1)
        f_ms, f_n, _ = ctx.last_fill_stats()  # fill-kernel device time of the LAST step (library CUDA events on this stream)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        assert int(d_status.item()) == 0, "device status != 0"
        return ms / steps, (ctx.launch_count - launches0) // steps, f_ms, f_n, clocks

    def gather_only_ms(want_cigar: bool):
        """Device time of the NCCL gather alone (same tensors as the timed step)."""
        if world == 1:
            return None
        for _ in range(2):
            shard.gather_device(d_score, d_off if want_cigar else None, d_cig if want_cigar else None, 16)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            shard.gather_device(d_score, d_off if want_cigar else None, d_cig if want_cigar else None, 16)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / 3], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    peak, peak_src = hbm_peak()
    n_chunks = -(-P // (1 << 18))  # chunk_pairs default: launches of the dominant kernel pair per step

    def ncu_measured(kernel):
        """DRAM bytes and warp instructions per 262,144-pair launch from the ncu capture of THIS build
        (profiles/ncu_traffic.json, regenerated by tools/ncu_traffic.py in the round's ncu step)."""
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                m = json.load(f)[kernel]
            m.setdefault("units_per_launch", 262144)
            return m
        except Exception:
            return None

    # ---- C3 (headline): traceback + CIGAR, dnaTwoBit inputs resident in HBM ---------------------------
    ms3, launches3, fill3, filln3, clocks3 = timed_device(True, True)
    g3 = cells_all / (ms3 * 1e-3) / 1e9
    alg3 = cells * BYTES_PER_CELL_TRACE + P * BYTES_PER_PAIR_SCORE  # this rank, per step
    ach3 = alg3 / (fill3 * 1e-3) / 1e9 if fill3 > 0 else None
    meas3 = ncu_measured("ckpt_path")
    line = {
        "metric": "GCUPS", "value": g3, "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "u16x2 (score pass) + int32 (path recompute)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "matrix": "HumanChimpTwo", "gap_open": GAP_OPEN, "gap_extend": GAP_EXTEND,
                   "pairs_total": total_pairs, "pairs_per_gpu": P, "cells_per_step_per_gpu": cells,
                   "inputs": "dnaTwoBit words resident in HBM (gnx_batch_device_twobit): packed 16-bit kernels stage them by "
                             "TMA, the screening and recompute kernels read the words as well",
                   "l2": "inputs (%.1f GB/GPU packed) exceed the 126 MB L2; no flush needed" % ((P * (WN + WM) * 8) / 1e9),
                   "sharding": "contiguous pair shards per rank; NCCL all_gather of scores, cigar counts and cigar records "
                               "inside the timed step (N>1)", "synth_seconds": round(gen_s, 1)},
        "gpu_launches": launches3,
        "clocks": clocks3,
        "roofline": {"bound": "hbm", "achieved": ach3, "peak": peak, "unit": "GB/s", "frac": (ach3 / peak) if ach3 else None,
                     "traffic": (meas3["dram_bytes"] * (P / n_chunks) / meas3["units_per_launch"]) if meas3 else None,
                     "algorithmic_bytes_per_launch": alg3 / n_chunks, "peak_source": peak_src,
                     "kernel": "affine_fill16_kernel<FREE,CM,CKPT,TB> + ckpt_classify_kernel + affine_ckpt_trace_kernel "
                               "(checkpoint-and-recompute: one launch of each per 262,144-pair chunk, traceback walk included)",
                     "fill_ms_per_step": fill3, "fill_launches_per_step": filln3, "algorithmic_bytes_per_step": alg3,
                     "note": "achieved = ALGORITHMIC bytes (0.75 B/cell of trace codes + 187 B/pair) / device time of the DP "
                             "kernels; the path keeps checkpoints instead of a trace matrix, so measured DRAM traffic is "
                             "below the algorithmic figure and the binding roof is integer issue (DESIGN.md section 6)"},
    }
    if world > 1:
        gms = gather_only_ms(True)
        line["gather"] = {"collective": "NCCL all_gather_into_tensor x5 (sizes, scores, counts, cigars)",
                          "bytes_received_per_rank": gather_bytes[0], "ms": gms,
                          "GBps_per_rank": gather_bytes[0] / (gms * 1e-3) / 1e9 if gms else None}

    # ---- the same step with byte-per-base inputs, and C2 (score only) both ways -----------------------
    ms3b, _, fill3b, _, _ = timed_device(True, False, steps=3)
    line["traceback_byte_inputs"] = {"value": cells_all / (ms3b * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": ms3b,
                                     "fill_ms_per_step": fill3b,
                                     "note": "C3 with []dna.Base bytes resident in HBM (gnx_batch_device)"}
    ms2, launches2, fill2, filln2, clocks2 = timed_device(False, True)
    ms2b, _, fill2b, _, _ = timed_device(False, False, steps=3)
    alg2 = P * BYTES_PER_PAIR_SCORE
    meas2 = ncu_measured("affine_fill16_kernel")
    so = {"workload": "C2 (configs[1]): same pairs, score only", "value": cells_all / (ms2 * 1e-3) / 1e9, "unit": "GCUPS",
          "ms_per_step": ms2, "gpu_launches": launches2, "clocks": clocks2,
          "byte_inputs": {"value": cells_all / (ms2b * 1e-3) / 1e9, "ms_per_step": ms2b, "fill_ms_per_step": fill2b},
          "roofline": {"bound": "hbm", "achieved": alg2 / (fill2 * 1e-3) / 1e9 if fill2 > 0 else None, "peak": peak,
                       "unit": "GB/s", "frac": alg2 / (fill2 * 1e-3) / 1e9 / peak if fill2 > 0 else None,
                       "traffic": (meas2["dram_bytes"] * (P / n_chunks) / meas2["units_per_launch"]) if meas2 else None,
                       "kernel": "affine_fill16_kernel<FREE,CM,TB> (packed 16-bit, 4 pairs/warp, 2-bit inputs by TMA)",
                       "fill_ms_per_step": fill2, "fill_launches_per_step": filln2, "algorithmic_bytes_per_step": alg2,
                       "note": "score-only moves 187 B per 75,000-cell pair: HBM cannot bind it (see issue_roofline)"}}
    if fill2 > 0 and clocks2.get("sm_mhz") and meas2 and meas2.get("inst_executed"):
        sm = torch.cuda.get_device_properties(local).multi_processor_count
        slots = sm * 4 * clocks2["sm_mhz"] * 1e6  # warp-instruction issue slots per second
        avail = slots / (cells / 64 / (fill2 * 1e-3))  # 32 lanes x 2 packed pairs = 64 cells per warp-instruction slot
        inst64 = meas2["inst_executed"] / (meas2["units_per_launch"] * N_LEN * M_LEN / 64)
        so["issue_roofline"] = {"bound": "issue", "issue_slots_per_64_cells": avail, "inst_per_64_cells": inst64,
                                "frac": inst64 / avail,
                                "note": "warp-instruction issue slots available per 64 DP cells at the measured fill rate vs "
                                        "the instructions the kernel executes for them (smsp__inst_executed of this build's "
                                        "ncu capture, profiles/ncu_traffic.json); frac = issue utilisation"}
    line["score_only"] = so

    # ---- other BASELINE shapes: C1, C4 (full per-GPU share, device-resident AND through the host API), const gap --
    if not args.no_extra:
        def run_shape(kind, n_len, m_len, pairs, want_cigar, cap_per_pair, steps=3, ctx=ctx, seed=SEED + 7):
            a, sao, b, sbo = synth_pairs(seed, pairs, n_len, m_len, first_pair=rank * pairs)
            ta, tb = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
            tao, tbo = torch.from_numpy(sao).to(dev), torch.from_numpy(sbo).to(dev)
            sc = torch.zeros(pairs, dtype=torch.int64, device=dev)
            off = torch.zeros(pairs + 1, dtype=torch.int64, device=dev)
            cg = torch.zeros(pairs * cap_per_pair * 16, dtype=torch.uint8, device=dev)

            def one():
                ctx.batch_device(kind, ta.data_ptr(), tao.data_ptr(), tb.data_ptr(), tbo.data_ptr(), sao, sbo, pairs, S,
                                 GAP_OPEN if kind != 2 else -430, GAP_EXTEND, want_cigar, sc.data_ptr(), cg.data_ptr(),
                                 off.data_ptr(), pairs * cap_per_pair, d_status.data_ptr(), stream)
            one()
            one()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                one()
            e1.record()
            barrier()
            ms_ = e0.elapsed_time(e1)
            f_ms, _, _ = ctx.last_fill_stats()
            if world > 1:
                t = torch.tensor([ms_], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms_ = float(t.item())
            assert int(d_status.item()) == 0
            return world * pairs * n_len * m_len * steps / (ms_ * 1e-3) / 1e9, ms_ / steps, f_ms, (a, sao, b, sbo, sc, off, cg)

        # ragged read batches (target 300-500 x query 100-150, trimmed copies of the C3 pairs): the packed 16-bit kernels
        # on host-binned quads -- what a batch of adapter-trimmed reads against variable windows looks like
        def run_ragged(pairs, steps=3):
            a, sao, b, sbo = synth_pairs(SEED + 9, pairs, N_LEN, M_LEN, first_pair=rank * pairs)
            rr = np.random.default_rng(SEED + 9 + rank)
            nl, ml = rr.integers(300, N_LEN + 1, size=pairs), rr.integers(100, M_LEN + 1, size=pairs)
            rao, rbo = np.zeros(pairs + 1, np.int64), np.zeros(pairs + 1, np.int64)
            np.cumsum(nl, out=rao[1:])
            np.cumsum(ml, out=rbo[1:])
            # keep the first nl bases of each target and the first ml of each query (vectorised gather)
            ia = np.repeat(sao[:-1] - rao[:-1], nl) + np.arange(rao[-1])
            ib = np.repeat(sbo[:-1] - rbo[:-1], ml) + np.arange(rbo[-1])
            ra, rb = a[ia], b[ib]
            ta, tb = torch.from_numpy(ra).to(dev), torch.from_numpy(rb).to(dev)
            tao, tbo = torch.from_numpy(rao).to(dev), torch.from_numpy(rbo).to(dev)
            sc = torch.zeros(pairs, dtype=torch.int64, device=dev)
            off = torch.zeros(pairs + 1, dtype=torch.int64, device=dev)
            cg = torch.zeros(pairs * 12 * 16, dtype=torch.uint8, device=dev)
            rcells = int((nl * ml).sum())
            res = {}
            for want in (True, False):
                def one():
                    ctx.batch_device(1, ta.data_ptr(), tao.data_ptr(), tb.data_ptr(), tbo.data_ptr(), rao, rbo, pairs, S, GAP_OPEN,
                                     GAP_EXTEND, want, sc.data_ptr(), cg.data_ptr(), off.data_ptr(), pairs * 12, d_status.data_ptr(), stream)
                one()
                one()
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    one()
                e1.record()
                barrier()
                ms_ = e0.elapsed_time(e1) / steps
                if world > 1:
                    t = torch.tensor([ms_], dtype=torch.float64, device=dev)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    ms_ = float(t.item())
                assert int(d_status.item()) == 0
                res["traceback" if want else "score_only"] = {"value": world * rcells / (ms_ * 1e-3) / 1e9, "unit": "GCUPS",
                                                               "ms_per_step": ms_, "kernel_path": list(ctx.last_kernel_path())}
            if rank == 0 and not args.no_cpu:
                import oracle as orc
                k = min(pairs, 4000)
                osc, ooff, ocig = orc.batch(ra[:rao[k]], rao[:k + 1], rb[:rbo[k]], rbo[:k + 1], orc.HUMAN_CHIMP_TWO_SCORE_MATRIX,
                                            GAP_OPEN, GAP_EXTEND, 1, False, os.cpu_count() or 1)
                res["parity_spot_check"] = bool(np.array_equal(osc, sc[:k].cpu().numpy()))
            res["pairs_per_gpu"] = pairs
            res["note"] = ("AffineGapLocal on RAGGED pairs, target 300-500 x query 100-150 (uniformly drawn), device-resident: "
                           "kernel_path (17, 1) / (16, 1) = the packed 16-bit kernels on host-binned quads")
            return res
        ragged = run_ragged(2_000_000)
        g1, ms1, _, _ = run_shape(0, 1000, 150, 100_000, True, 16)
        gc, msc, _, _ = run_shape(2, N_LEN, M_LEN, 1_000_000, True, 400)
        # C4 (configs[3]): 100k pairs of 10 kb x 10 kb over 8 GPUs = 12,500 per GPU, global affine + CIGAR
        p4 = args.c4_pairs
        g4, ms4, f4, keep = run_shape(0, 10_000, 10_000, p4, True, 600, steps=2, seed=20260104)
        a4, ao4, b4, bo4, sc4, off4, cg4 = keep
        cells4 = p4 * 10_000 * 10_000
        alg4 = cells4 * BYTES_PER_CELL_TRACE
        meas4 = ncu_measured("affine_long_kernel")
        c4 = {"value": g4, "unit": "GCUPS", "pairs_per_gpu": p4, "ms_per_step": ms4,
              "roofline": {"bound": "hbm", "achieved": alg4 / (f4 * 1e-3) / 1e9 if f4 > 0 else None, "peak": peak, "unit": "GB/s",
                           "frac": alg4 / (f4 * 1e-3) / 1e9 / peak if f4 > 0 else None,
                           "traffic": (meas4["dram_bytes"] * p4 / meas4["units_per_launch"]) if meas4 else None,
                           "algorithmic_bytes_per_launch": alg4, "kernel": "affine_long_kernel<FREE=0> (one launch per step)",
                           "fill_ms_per_step": f4, "peak_source": peak_src},
              "note": "AffineGap (global) + CIGAR, device-resident; tile checkpoints + recompute of the route's tiles in "
                      "per-warp scratch (gnx_long.cuh), no trace matrix"}
        # through the host-buffer API (pageable numpy inputs and outputs: what a Go caller holds), H2D + D2H inside
        t0 = time.perf_counter()
        hsc, hoff, hcig = ctx.affine_gap_batch(a4, ao4, b4, bo4, S, GAP_OPEN, GAP_EXTEND, False, True, cigar_cap=p4 * 600)
        dt4 = time.perf_counter() - t0
        t0 = time.perf_counter()
        hsc, hoff, hcig = ctx.affine_gap_batch(a4, ao4, b4, bo4, S, GAP_OPEN, GAP_EXTEND, False, True, cigar_cap=p4 * 600)
        dt4 = min(dt4, time.perf_counter() - t0)
        if world > 1:
            t = torch.tensor([dt4], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt4 = float(t.item())
        c4["e2e"] = {"value": world * cells4 / dt4 / 1e9, "unit": "GCUPS", "seconds": dt4,
                     "h2d_bytes_per_step": int(len(a4) + len(b4) + 16 * (p4 + 1)),
                     "d2h_bytes_per_step": int(8 * p4 + 8 * (p4 + 1) + 16 * int(hoff[-1])),
                     "api": "gnx_affine_batch, pageable host buffers"}
        # parity: host API == device API everywhere; size-independent properties on every pair; oracle on a few
        ok4 = bool(np.array_equal(hsc, sc4.cpu().numpy()) and np.array_equal(hoff, off4.cpu().numpy()))
        rl, op = hcig["run_length"], hcig["op"]
        ok4 &= bool(np.all(np.add.reduceat(np.where(op != 1, rl, 0), hoff[:-1]) == 10_000)
                    and np.all(np.add.reduceat(np.where(op != 2, rl, 0), hoff[:-1]) == 10_000))
        if rank == 0 and not args.no_cpu:
            import oracle as orc
            k4 = min(p4, 2 * (os.cpu_count() or 1))
            osc, ooff, ocig = orc.batch(a4[:ao4[k4]], ao4[:k4 + 1], b4[:bo4[k4]], bo4[:k4 + 1], orc.HUMAN_CHIMP_TWO_SCORE_MATRIX,
                                        GAP_OPEN, GAP_EXTEND, 0, True, os.cpu_count() or 1)
            t4 = int(ooff[-1])
            ok4 &= bool(np.array_equal(hsc[:k4], osc) and np.array_equal(hoff[:k4 + 1], ooff)
                        and np.array_equal(rl[:t4], ocig["run_length"]) and np.array_equal(op[:t4], ocig["op"]))
            c4["parity_pairs_vs_oracle"] = k4
        c4["parity_spot_check"] = ok4
        del a4, b4, sc4, off4, cg4, keep, hsc, hoff, hcig
        torch.cuda.empty_cache()
        line["other_workloads"] = {
            "c3_ragged_300-500x100-150": ragged,
            "c1_global_1000x150_traceback": {"value": g1, "unit": "GCUPS", "pairs_per_gpu": 100_000, "ms_per_step": ms1,
                                             "note": "AffineGap (global) + CIGAR, configs[0] shape x100"},
            "c4_global_10kx10k_traceback": c4,
            "const_gap_500x150_traceback": {"value": gc, "unit": "GCUPS", "pairs_per_gpu": 1_000_000,
                                            "ms_per_step": msc, "note": "ConstGap_highMem + CIGAR, g=-430"}}
        # 2-bit packing (HBM-bound) and the perfect-match seed / extend steps
        line["other_workloads"].update(twobit_block(ctx, L, dev, stream, rank, world, barrier, args))
        # SURVEY 8f-3 / configs[4]: the whole per-read gsw driver on paired reads against a >= 1 Gb reference
        line["other_workloads"]["c5_gsw_paired_2x150"] = gsw_block(ctx, L, dev, rank, world, args)

    # ---- e2e: the public host-buffer API with H2D + D2H inside the timed region ------------------------
    # Headline e2e = what the cgo shim of INTEGRATION.md does: PAGEABLE []dna.Base bytes in, pageable results out.
    # The other variants show what the same call delivers with page-locked buffers (gnx_host_alloc) and with the
    # batch in dnaTwoBit form (a quarter of the H2D bytes).
    if not args.no_e2e:
        e2e_steps = max(2, min(args.steps, 5))
        p_score, p_off, p_cig = pinned_array(P, np.int64), pinned_array(P + 1, np.int64), pinned_array(cig_cap, CIGAR_DTYPE)
        g_alpha, g_beta = np.array(h_alpha), np.array(h_beta)          # pageable copies ("Go slices")
        g_wa, g_wb = np.array(h_wa), np.array(h_wb)
        g_score, g_off, g_cig = np.zeros(P, np.int64), np.zeros(P + 1, np.int64), np.zeros(cig_cap, CIGAR_DTYPE)

        def e2e(want_cigar: bool, twobit: bool, pinned: bool, pack: bool = True):
            # pack: the library packs pageable bytes to 2 bits per base while it stages them (its default)
            ctx.set_option("pack_stage", 1 if pack else 0)
            out = (p_score, p_off if want_cigar else None, p_cig if want_cigar else None) if pinned else \
                  (g_score, g_off if want_cigar else None, g_cig if want_cigar else None)

            def call():
                if twobit:
                    wa, wb = (h_wa, h_wb) if pinned else (g_wa, g_wb)
                    ctx.affine_gap_batch_twobit(wa, N_LEN, wb, M_LEN, S, GAP_OPEN, GAP_EXTEND, True, want_cigar, out=out, n_pairs=P)
                else:
                    al, be = (h_alpha, h_beta) if pinned else (g_alpha, g_beta)
                    ctx.affine_gap_batch(al, ao, be, bo, S, GAP_OPEN, GAP_EXTEND, True, want_cigar, out=out)
            call()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                call()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / e2e_steps
            if world > 1:
                t = torch.tensor([dt], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            h2d = (P * (WN + WM) * 8) if twobit else (na + nb + 2 * (P + 1) * 8)  # the host buffers handed to the call
            packed = twobit or pack
            pcie = (P * (WN + WM) * 8) if packed else (na + nb + 2 * (P + 1) * 8)  # what crosses PCIe
            d2h = P * 8 + ((P + 1) * 8 + int(out[1][-1]) * 16 if want_cigar else 0)
            return {"value": cells_all / dt / 1e9, "unit": "GCUPS", "ms_per_step": dt * 1e3, "h2d_bytes_per_step": int(h2d),
                    "pcie_h2d_bytes_per_step": int(pcie), "d2h_bytes_per_step": int(d2h)}, out

        head, out3 = e2e(True, False, False)
        head["api"] = ("gnx_affine_batch: pageable []dna.Base bytes in, pageable scores + cigars out (the cgo shim's call); "
                       "the library's staging pass packs the bytes to dnaTwoBit words on the host threads")
        assert np.array_equal(out3[0], d_score.cpu().numpy()), "host-API scores differ from device-API scores"
        if rank == 0 and not args.no_cpu:
            # the first pairs of the timed batch diffed against the oracle, score AND cigar
            import oracle as orc
            k = min(P, 20000)
            osc, ooff, ocig = orc.batch(h_alpha[:k * N_LEN], ao[:k + 1], h_beta[:k * M_LEN], bo[:k + 1],
                                        orc.HUMAN_CHIMP_TWO_SCORE_MATRIX, GAP_OPEN, GAP_EXTEND, 1, True, os.cpu_count() or 1)
            t = int(ooff[-1])
            line["parity_spot_check"] = bool(
                np.array_equal(out3[0][:k], osc) and np.array_equal(out3[1][:k + 1], ooff)
                and np.array_equal(out3[2]["run_length"][:t], ocig["run_length"]) and np.array_equal(out3[2]["op"][:t], ocig["op"]))
        ref_cig = (out3[1].copy(), out3[2][:int(out3[1][-1])].copy())
        variants = {}
        for name, tb_, pin_, pk_ in (("pageable_bytes_staged_unpacked", False, False, False), ("pinned_bytes", False, True, True),
                                     ("pageable_twobit", True, False, True), ("pinned_twobit", True, True, True)):
            variants[name], o = e2e(True, tb_, pin_, pk_)
            tot = int(o[1][-1])
            assert np.array_equal(o[1], ref_cig[0]) and np.array_equal(o[2]["run_length"][:tot], ref_cig[1]["run_length"]) \
                and np.array_equal(o[2]["op"][:tot], ref_cig[1]["op"]), f"e2e variant {name} differs"
        head["variants"] = variants
        line["e2e"] = head
        so_e2e = {}
        for name, tb_, pin_, pk_ in (("pageable_bytes", False, False, True), ("pageable_bytes_staged_unpacked", False, False, False),
                                     ("pinned_bytes", False, True, True), ("pageable_twobit", True, False, True),
                                     ("pinned_twobit", True, True, True)):
            so_e2e[name], o = e2e(False, tb_, pin_, pk_)
            assert np.array_equal(o[0], d_score.cpu().numpy()), f"score-only e2e variant {name} differs"
        line["score_only"]["e2e"] = so_e2e
        ctx.set_option("pack_stage", 1)
        del g_alpha, g_beta, g_wa, g_wb, g_score, g_off, g_cig

    # ---- CPU baseline (rank 0, N=1 only): the oracle port on the host cores, same workload (C3) --------
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        sample = max(2000, 12000 * threads)  # ~10 s of host work
        g, dt = cpu_reference_gcups(sample, True, threads)
        line["cpu_baseline"] = {"value": g, "unit": "GCUPS", "cores": threads, "kind": "port",
                                "sample": f"first {sample} pairs of the batch, traceback + cigar, {dt:.1f} s "
                                          "(C restatement of the Go path; Go toolchain absent)"}

    if rank == 0:
        emit(line)
    ctx.close()
    for ptr in pinned_ptrs:
        L.gnx_host_free(ptr)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
