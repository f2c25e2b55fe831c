#!/usr/bin/env python
"""bench.py -- GCUPS of the affine-gap DP hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--pairs P] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic pairs resident in HBM.
Workload (BASELINE.json configs[1], "C2"): P = 10^7 pairs per GPU, target 500 x query 150,
AffineGapLocal semantics (free end gaps), HumanChimpTwo matrix, O=-600, E=-150, score only.
The same line also carries configs[2] ("C3": the same pairs with full traceback + CIGAR) under
"traceback".  GCUPS counts each DP cell once: sum(n*m) / seconds / 1e9.
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
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
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
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_gcups(n_pairs: int, want_cigar: bool, threads: int):
    """The oracle (C restatement of the Go path) on the host cores: the reported CPU baseline."""
    import oracle as orc
    from gonomics_b200.synth import synth_pairs
    a, ao, b, bo = synth_pairs(SEED, n_pairs, N_LEN, M_LEN)
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
        g, dt = cpu_reference_gcups(sample, False, threads)
        if i >= args.warmup:
            vals.append((g, dt))
    gcups = sample * N_LEN * M_LEN * len(vals) / sum(d for _, d in vals) / 1e9
    line = {
        "impl": "reference", "metric": "GCUPS", "value": gcups, "unit": "GCUPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(d for _, d in vals) / len(vals),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": "C2: target 500 x query 150 semi-global affine (AffineGapLocal), score only",
                   "matrix": "HumanChimpTwo", "gap_open": GAP_OPEN, "gap_extend": GAP_EXTEND},
        "cpu_baseline": {"value": gcups, "unit": "GCUPS", "cores": threads, "kind": "port",
                         "sample": f"{sample} pairs of the C2 batch per step (C restatement of the Go path; "
                                   "Go toolchain absent)"},
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
    ap.add_argument("--pairs", type=int, default=10_000_000, help="pairs per GPU per step")
    ap.add_argument("--impl", default="gnx", choices=["gnx", "reference"])
    ap.add_argument("--no-traceback", action="store_true", help="skip the C3 (traceback) block")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the C1 / C4-sample / const-gap block")
    ap.add_argument("--quick", action="store_true", help="profiling runs: no e2e/cpu legs, warm-up not clamped")
    args = ap.parse_args()
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
        args.no_e2e = args.no_cpu = True
    else:
        args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from gonomics_b200 import align
    from gonomics_b200._lib import CIGAR_DTYPE, load
    from gonomics_b200.synth import synth_pairs

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    P = args.pairs
    S = align.HumanChimpTwoScoreMatrix
    L = load()
    ctx = align.Context(local)
    cells = P * N_LEN * M_LEN

    # ---- synthetic inputs: this rank's shard of the global batch, in pinned host memory ----------
    na, nb = P * N_LEN, P * M_LEN
    import ctypes as C
    pa, pb_ = L.gnx_host_alloc(na), L.gnx_host_alloc(nb)
    h_alpha = np.ctypeslib.as_array(C.cast(pa, C.POINTER(C.c_uint8)), shape=(na,))
    h_beta = np.ctypeslib.as_array(C.cast(pb_, C.POINTER(C.c_uint8)), shape=(nb,))
    t0 = time.perf_counter()
    _, ao, _, bo = synth_pairs(SEED, P, N_LEN, M_LEN, first_pair=rank * P, alpha_out=h_alpha, beta_out=h_beta)
    ao -= ao[0]
    bo -= bo[0]
    gen_s = time.perf_counter() - t0
    d_alpha = torch.from_numpy(h_alpha).to(dev)
    d_beta = torch.from_numpy(h_beta).to(dev)
    d_ao, d_bo = torch.from_numpy(ao).to(dev), torch.from_numpy(bo).to(dev)
    d_score = torch.zeros(P, dtype=torch.int64, device=dev)
    d_status = torch.zeros(1, dtype=torch.int32, device=dev)
    cig_cap = P * 12
    d_cig = torch.zeros(cig_cap * 16, dtype=torch.uint8, device=dev)
    d_off = torch.zeros(P + 1, dtype=torch.int64, device=dev)
    gathered = torch.zeros(world * P, dtype=torch.int64, device=dev) if world > 1 else None
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step(want_cigar: bool):
        ctx.batch_device(1, d_alpha.data_ptr(), d_ao.data_ptr(), d_beta.data_ptr(), d_bo.data_ptr(), ao, bo, P, S,
                         GAP_OPEN, GAP_EXTEND, want_cigar, d_score.data_ptr(), d_cig.data_ptr() if want_cigar else 0,
                         d_off.data_ptr() if want_cigar else 0, cig_cap, d_status.data_ptr(), stream)
        if world > 1:  # the path's only exchange: gather the per-shard scores (north_star)
            dist.all_gather_into_tensor(gathered, d_score)

    def timed_device(want_cigar: bool):
        for _ in range(args.warmup):
            device_step(want_cigar)
        barrier()
        launches0 = ctx.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fill_ms = 0.0
        fill_launches = 0
        sampler = ClockSampler(local)
        sampler.start()
        e0.record()
        for _ in range(args.steps):
            device_step(want_cigar)
        e1.record()
        barrier()
        clocks = sampler.stop()
        ms = e0.elapsed_time(e1)
        # fill-kernel device time of the LAST step (CUDA events recorded by the library on this stream)
        f_ms, f_n, _ = ctx.last_fill_stats()
        fill_ms, fill_launches = f_ms, f_n
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        assert int(d_status.item()) == 0, "device status != 0"
        return ms, ctx.launch_count - launches0, fill_ms, fill_launches, clocks

    peak, peak_src = hbm_peak()

    # ---- C2: score only (headline value) ---------------------------------------------------------
    ms, launches, fill_ms, fill_n, clocks = timed_device(False)
    gcups = world * cells * args.steps / (ms * 1e-3) / 1e9
    alg_bytes = P * BYTES_PER_PAIR_SCORE  # per step, all fill launches of the step together
    achieved = alg_bytes / (fill_ms * 1e-3) / 1e9 if fill_ms > 0 else None
    line = {
        "metric": "GCUPS", "value": gcups, "unit": "GCUPS", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u16x2", "data": "synthetic",
        "config": {"workload": "C2 (BASELINE.json configs[1]): %d pairs/GPU, target 500 x query 150, semi-global "
                               "affine gap (AffineGapLocal), score only" % P,
                   "matrix": "HumanChimpTwo", "gap_open": GAP_OPEN, "gap_extend": GAP_EXTEND,
                   "pairs_per_gpu": P, "cells_per_step_per_gpu": cells,
                   "l2": "inputs (%.1f GB/GPU) exceed the 126 MB L2; no flush needed" % ((na + nb) / 1e9),
                   "sharding": "independent pair shards per rank; all_gather of scores only (N>1)",
                   "synth_seconds": round(gen_s, 1)},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None,
                     "traffic": ncu_traffic("affine_fill16_kernel", P / max(fill_n, 1)),
                     "algorithmic_bytes_per_launch": alg_bytes / max(fill_n, 1),
                     "peak_source": peak_src, "kernel": "affine_fill16_kernel<FREE=1> (packed 16-bit, 4 pairs/warp)",
                     "fill_ms_per_step": fill_ms, "fill_launches_per_step": fill_n,
                     "algorithmic_bytes_per_step": alg_bytes,
                     "note": "score-only fill moves 187 B per 75,000-cell pair: HBM cannot bind it; the binding "
                             "roof is integer issue (see issue_roofline and DESIGN.md)"},
    }
    if fill_ms > 0 and clocks.get("sm_mhz"):
        sm = torch.cuda.get_device_properties(local).multi_processor_count
        slots = sm * 4 * clocks["sm_mhz"] * 1e6  # warp-instruction issue slots per second
        # 32 lanes x 2 packed pairs = 64 cells per warp-instruction slot in the packed kernel
        avail = slots / (cells / 64 / (fill_ms * 1e-3))
        # instructions the kernel executes per 64 cells: smsp__inst_executed.sum of the ncu capture in
        # profiles/r01i_ckpt_path.md (3.0786e9 for a 262,144-pair chunk of 75,000-cell pairs)
        inst64 = 3.0786e9 / (262144 * 75000 / 64)
        line["issue_roofline"] = {"bound": "issue", "cells_per_s_fill": cells / (fill_ms * 1e-3),
                                  "issue_slots_per_s": slots, "issue_slots_per_64_cells": avail,
                                  "inst_per_64_cells": inst64, "frac": inst64 / avail,
                                  "steady_loop_inst_per_64_cells": 8.6,
                                  "note": "the binding roof of this integer max-plus kernel: warp-instruction issue "
                                          "slots available per 64 DP cells (one packed warp-cell) at the measured fill "
                                          "rate vs the instructions the kernel executes for them (ncu); frac = issue "
                                          "utilisation.  The steady loop itself needs 8.6 (7 per packed cell + per-step "
                                          "overhead), see DESIGN.md"}

    # ---- C3: traceback + CIGAR on the same pairs -------------------------------------------------
    if not args.no_traceback:
        ms3, launches3, fill3, filln3, clocks3 = timed_device(True)
        g3 = world * cells * args.steps / (ms3 * 1e-3) / 1e9
        alg3 = cells * BYTES_PER_CELL_TRACE + P * BYTES_PER_PAIR_SCORE
        n_chunks3 = -(-P // (1 << 18))  # chunk_pairs default
        ckpt_path = filln3 >= 2 * n_chunks3  # fill16+checkpoints and the recompute kernel: two fill launches per chunk
        ach3 = alg3 / (fill3 * 1e-3) / 1e9 if fill3 > 0 else None
        line["traceback"] = {
            "workload": "C3 (configs[2]): same pairs, full traceback + CIGAR", "value": g3, "unit": "GCUPS",
            "ms_per_step": ms3 / args.steps, "gpu_launches": launches3, "clocks": clocks3,
            "roofline": {"bound": "hbm", "achieved": ach3, "peak": peak, "unit": "GB/s",
                         "frac": (ach3 / peak) if ach3 else None,
                         "traffic": ncu_traffic("ckpt_path" if ckpt_path else "affine_fill3_kernel_trace",
                                                P / max(n_chunks3, 1)),
                         "algorithmic_bytes_per_launch": alg3 / max(n_chunks3, 1), "peak_source": peak_src,
                         "kernel": ("affine_fill16_kernel<FREE,CM,CKPT> + affine_ckpt_trace_kernel (checkpoint-and-"
                                    "recompute: two launches per chunk, traceback walk included)") if ckpt_path
                         else "affine_fill3_kernel<C=10,LPP=16,MODE=2,FREE=1>", "fill_ms_per_step": fill3,
                         "fill_launches_per_step": filln3, "algorithmic_bytes_per_step": alg3}}

    # ---- other BASELINE shapes (device-resident, fewer steps): C1, a C4-shaped sample, constant gap ----
    if not args.quick and not args.no_extra:
        def run_shape(kind, n_len, m_len, pairs, want_cigar, cap_per_pair, steps=3, ctx=ctx):
            a, sao, b, sbo = synth_pairs(SEED + 7, pairs, n_len, m_len, first_pair=rank * pairs)
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
            if world > 1:
                t = torch.tensor([ms_], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms_ = float(t.item())
            assert int(d_status.item()) == 0
            return world * pairs * n_len * m_len * steps / (ms_ * 1e-3) / 1e9, ms_ / steps

        g1, ms1 = run_shape(0, 1000, 150, 100_000, True, 16)
        # C4: a 10 kb x 10 kb pair owns 82 MB of traceback matrix and the one-warp-per-pair kernel needs pairs in
        # flight to fill the SMs, so this block gets its own context sized for the 180 GB part (75 % of what is free)
        free_b, _ = torch.cuda.mem_get_info(dev)
        ws4 = int(free_b * 0.75)
        pairs4 = max(256, min(1776, ws4 // 82_300_000) // 4 * 4)
        ctx4 = align.Context(local, ws4)
        try:
            g4, ms4 = run_shape(0, 10_000, 10_000, pairs4, True, 4096, steps=2, ctx=ctx4)
        finally:
            ctx4.close()
        torch.cuda.empty_cache()
        gc, msc = run_shape(2, N_LEN, M_LEN, 1_000_000, True, 400)
        line["other_workloads"] = {
            "c1_global_1000x150_traceback": {"value": g1, "unit": "GCUPS", "pairs_per_gpu": 100_000, "ms_per_step": ms1,
                                             "note": "AffineGap (global) + CIGAR, configs[0] shape x100"},
            "c4_global_10kx10k_traceback": {"value": g4, "unit": "GCUPS", "pairs_per_gpu": pairs4, "ms_per_step": ms4,
                                            "workspace_gb": round(ws4 / 1e9, 1),
                                            "note": "AffineGap (global) + CIGAR, one warp per pair through 32 strips "
                                                    "(one workspace-sized chunk of configs[3]: 82 MB of traceback "
                                                    "matrix per pair, dedicated context with 75 % of free HBM)"},
            "const_gap_500x150_traceback": {"value": gc, "unit": "GCUPS", "pairs_per_gpu": 1_000_000,
                                            "ms_per_step": msc, "note": "ConstGap_highMem + CIGAR, g=-430"}}

    # ---- 2-bit packing (HBM-bound) and the perfect-match seed step ----------------------
    if not args.quick and not args.no_extra:
        line["other_workloads"].update(twobit_block(ctx, L, dev, stream, rank, world, barrier, args))

    # ---- e2e: the public host-buffer API, pinned host inputs, H2D + D2H inside the timed region --
    if not args.no_e2e:
        pinned = []

        def pinned_array(count, dtype):
            """Caller-owned result buffer in page-locked memory (gnx_host_alloc): results are DMA'd straight into it."""
            dt = np.dtype(dtype)
            ptr = L.gnx_host_alloc(max(count * dt.itemsize, 1))
            if not ptr:
                raise MemoryError("gnx_host_alloc failed")
            pinned.append(ptr)
            raw = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(max(count * dt.itemsize, 1),))
            return raw[:count * dt.itemsize].view(dt)

        def e2e(want_cigar: bool):
            out_score = pinned_array(P, np.int64)
            out_off = pinned_array(P + 1, np.int64) if want_cigar else None
            out_cig = pinned_array(cig_cap, CIGAR_DTYPE) if want_cigar else None
            out = (out_score, out_off, out_cig)
            for _ in range(2):
                ctx.affine_gap_batch(h_alpha, ao, h_beta, bo, S, GAP_OPEN, GAP_EXTEND, True, want_cigar, out=out)
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                ctx.affine_gap_batch(h_alpha, ao, h_beta, bo, S, GAP_OPEN, GAP_EXTEND, True, want_cigar, out=out)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            h2d = na + nb + 2 * (P + 1) * 8
            d2h = P * 8 + ((P + 1) * 8 + int(out_off[-1]) * 16 if want_cigar else 0)
            return world * cells * args.steps / dt / 1e9, h2d, d2h, (out_score if not want_cigar else out)
        v, h2d, d2h, sc_host = e2e(False)
        line["e2e"] = {"value": v, "unit": "GCUPS", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "api": "gnx_affine_batch (host buffers, inputs and results pinned), score only"}
        assert np.array_equal(sc_host, d_score.cpu().numpy()), "host-API scores differ from device-API scores"
        if not args.no_traceback:
            v3, h2d3, d2h3, out3 = e2e(True)
            line["traceback"]["e2e"] = {"value": v3, "unit": "GCUPS", "h2d_bytes_per_step": h2d3,
                                        "d2h_bytes_per_step": d2h3}
            if rank == 0 and world == 1 and not args.no_cpu:
                # the first pairs of the timed batch diffed against the oracle, score AND cigar
                import oracle as orc
                k = min(P, 20000)
                osc, ooff, ocig = orc.batch(h_alpha[:k * N_LEN], ao[:k + 1], h_beta[:k * M_LEN], bo[:k + 1],
                                            orc.HUMAN_CHIMP_TWO_SCORE_MATRIX, GAP_OPEN, GAP_EXTEND, 1, True,
                                            os.cpu_count() or 1)
                gsc, goff, gcig = out3
                t = int(ooff[-1])
                line["traceback"]["parity_spot_check"] = bool(
                    np.array_equal(gsc[:k], osc) and np.array_equal(goff[:k + 1], ooff)
                    and np.array_equal(gcig["run_length"][:t], ocig["run_length"])
                    and np.array_equal(gcig["op"][:t], ocig["op"]))
            out3 = None
        sc_host = None
        for ptr in pinned:
            L.gnx_host_free(ptr)

    # ---- CPU baseline (rank 0, N=1 only): the oracle port on the host cores ----------------------
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        sample = max(2000, 12000 * threads)  # ~10 s of host work
        g, dt = cpu_reference_gcups(sample, True, threads)
        line["cpu_baseline"] = {"value": g, "unit": "GCUPS", "cores": threads, "kind": "port",
                                "sample": f"first {sample} pairs of the batch, traceback + cigar, {dt:.1f} s "
                                          "(C restatement of the Go path; Go toolchain absent)"}
        # spot-check: the GPU scores of that prefix equal the oracle's
        import oracle as orc
        k = min(sample, 20000)
        osc, _, _ = orc.batch(h_alpha[:k * N_LEN], ao[:k + 1], h_beta[:k * M_LEN], bo[:k + 1],
                              orc.HUMAN_CHIMP_TWO_SCORE_MATRIX, GAP_OPEN, GAP_EXTEND, 1, False, threads)
        line["parity_spot_check"] = bool(np.array_equal(osc, d_score[:k].cpu().numpy()))

    if rank == 0:
        emit(line)
    ctx.close()
    L.gnx_host_free(pa)
    L.gnx_host_free(pb_)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
