"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the reference's golden
vectors.  Bit-exact: identical int64 score and identical (RunLength, Op) sequence."""
import numpy as np
import pytest

import oracle as orc
from golden_util import MATRICES, bases, cigar_to_beds, load, random_pair
from gonomics_b200 import _lib, align
from gonomics_b200.synth import synth_pairs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = align.Context(0)
    yield c
    c.close()


def concat(seqs):
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum([len(s) for s in seqs], out=off[1:])
    cat = np.concatenate([np.asarray(s, dtype=np.uint8) for s in seqs] + [np.zeros(0, dtype=np.uint8)])
    return cat, off


def check_batch(ctx, alphas, betas, S, O, E, mode, want_cigar=True, threads=8):
    """mode 0 global affine, 1 free-end affine, 2 const gap (O = penalty)."""
    ac, ao = concat(alphas)
    bc, bo = concat(betas)
    if mode == 2:
        sc, off, cig = ctx.const_gap_batch(ac, ao, bc, bo, S, O, want_cigar)
    else:
        sc, off, cig = ctx.affine_gap_batch(ac, ao, bc, bo, S, O, E, mode == 1, want_cigar)
    osc, ooff, ocig = orc.batch(ac, ao, bc, bo, S, O, E, mode, want_cigar, threads)
    bad = np.nonzero(sc != osc)[0]
    assert len(bad) == 0, f"score mismatch at pairs {bad[:5]}: gpu {sc[bad[:5]]} oracle {osc[bad[:5]]} " \
                          f"lens {[(len(alphas[i]), len(betas[i])) for i in bad[:5]]}"
    if want_cigar:
        if not (np.array_equal(off, ooff) and np.array_equal(cig["run_length"], ocig["run_length"])
                and np.array_equal(cig["op"], ocig["op"])):
            for p in range(len(alphas)):
                g = [(int(r), int(o)) for r, o in cig[off[p]:off[p + 1]]]
                w = [(int(r), int(o)) for r, o in ocig[ooff[p]:ooff[p + 1]]]
                assert g == w, f"cigar mismatch pair {p} (n={len(alphas[p])}, m={len(betas[p])}): " \
                               f"gpu {orc.print_cigar(g)} oracle {orc.print_cigar(w)}"
    return sc


# ---- the reference's golden vectors through the CUDA path ------------------------------------
def test_golden_affine_global(ctx):
    g = load("affine_global")
    S = MATRICES[g["matrix"]]
    for c in g["cases"]:
        a, b = bases(c["alpha"]), bases(c["beta"])
        _, cig = align.AffineGap_highMem(a, b, S, g["gap_open"], g["gap_extend"], ctx)
        assert align.View(a, b, cig) == c["view"]
        _, cig2 = align.AffineGap(a, b, S, g["gap_open"], g["gap_extend"], ctx)
        # TestAffineGap_lowMem (affineGap_test.go:57-81) asserts lowMem == highMem on these vectors with a 3 x 3 checker;
        # the multi-board driver itself is not reproduced: without the explicit opt-in the call is an error
        _, cig3 = align.AffineGap_customizeCheckersize(a, b, S, g["gap_open"], g["gap_extend"], 3, 3, ctx,
                                                       multi_board="highmem")
        assert cig2 == cig and cig3 == cig
        if len(a) > 3 or len(b) > 3:
            with pytest.raises(_lib.GnxError):
                align.AffineGap_customizeCheckersize(a, b, S, g["gap_open"], g["gap_extend"], 3, 3, ctx)


def test_golden_affine_local(ctx):
    for c in load("affine_local")["cases"]:
        score, cig = align.AffineGapLocal(bases(c["target"]), bases(c["query"]), MATRICES[c["matrix"]],
                                          c["gap_open"], c["gap_extend"], ctx)
        assert (score, align.PrintCigar(cig)) == (c["score"], c["cigar"])


def test_golden_const_gap(ctx):
    g = load("const_gap")
    S = MATRICES[g["matrix"]]
    for c in g["cases"]:
        a, b = bases(c["alpha"]), bases(c["beta"])
        _, cig = align.ConstGap(a, b, S, g["gap_pen"], ctx)
        assert align.View(a, b, cig) == c["view"]
    g = load("global_alignment")
    a, b = bases(g["alpha"]), bases(g["beta"])
    score, cig = align.ConstGap(a, b, MATRICES[g["matrix"]], g["gap_pen"], ctx)
    assert (score, align.PrintCigar(cig), align.View(a, b, cig)) == (-730, "3M3D3M", g["view"])


def test_golden_anchor(ctx):
    g = load("anchor")
    S = MATRICES[g["matrix"]]
    for c in g["cases"]:
        a, b = bases(c["alpha"], upper=True), bases(c["beta"], upper=True)
        score, cig = align.AffineGap_customizeCheckersize(a, b, S, g["gap_open"], g["gap_extend"], 10000, 10000, ctx)
        assert score == c["score"] and [list(x) for x in cig] == c["cigar"], c["region1"]


def test_golden_cigar_to_bed_10kb(ctx):
    g = load("cigar_to_bed")
    S = MATRICES[g["matrix"]]
    for c in g["cases"]:
        a, b = bases(c["alpha"], upper=True), bases(c["beta"], upper=True)
        score, cig = align.AffineGap(a, b, S, g["gap_open"], g["gap_extend"], ctx)
        ins, dele = cigar_to_beds(cig, c["first_pos_ins"], c["first_pos_del"], c["chrom"])
        assert ins == c["ins_bed"] and dele == c["del_bed"], c["files"]
        if len(a) > 9000:
            assert score == 790738 and len(cig) == 19
        assert (score, [tuple(x) for x in cig]) == orc.affine_gap_highmem(a, b, S, g["gap_open"], g["gap_extend"])


def test_golden_affine_chunk(ctx):
    g = load("affine_chunk")  # align/affineGap_test.go:83-93 TestAffineGapChunk
    S = MATRICES[g["matrix"]]
    for c in g["cases"]:
        a, b = bases(c["alpha"]), bases(c["beta"])
        score, cig = align.AffineGapChunk(a, b, S, g["gap_open"], g["gap_extend"], g["chunk"], ctx)
        assert align.View(a, b, cig) == c["view"]
        assert (score, [tuple(x) for x in cig]) == orc.affine_gap_chunk(a, b, S, g["gap_open"], g["gap_extend"], g["chunk"])
    with pytest.raises(_lib.GnxError) as ei:  # ragged length: the reference log.Fatalf's
        align.AffineGapChunk(bases("ACGT"), bases("ACG"), S, -400, -30, 3, ctx)
    assert ei.value.code == _lib.GNX_ECHUNK


def test_random_affine_chunk(ctx):
    rng = np.random.default_rng(800)
    S = orc.HUMAN_CHIMP_TWO_SCORE_MATRIX
    for chunk in (1, 2, 3, 7):
        al, be = [], []
        for _ in range(120):
            n, m = int(rng.integers(0, 60)) * chunk, int(rng.integers(0, 60)) * chunk
            unit = rng.integers(0, 4, chunk, dtype=np.uint8)
            a, b = np.resize(unit, n).copy(), np.resize(unit, m).copy()  # tandem repeats of a chunk-sized unit
            for arr in (a, b):
                if len(arr):
                    k = max(1, len(arr) // 15)
                    arr[rng.integers(0, len(arr), k)] = rng.integers(0, 4, k)
            al.append(a)
            be.append(b)
        ac, ao = concat(al)
        bc, bo = concat(be)
        sc, off, cig = ctx.affine_gap_chunk_batch(ac, ao, bc, bo, S, -600, -150, chunk)
        for p in range(len(al)):
            want = orc.affine_gap_chunk(al[p], be[p], S, -600, -150, chunk)
            got = (int(sc[p]), [(int(r), int(o)) for r, o in cig[off[p]:off[p + 1]]])
            assert got == want, (chunk, p, len(al[p]), len(be[p]))


def test_engine_fifo(ctx):
    cases = load("affine_local")["cases"][:4]
    inputs, outputs = align.GoAffineGapLocalEngine(MATRICES["Default"], -600, -150)
    for c in cases:  # align/affineGap_test.go:157-192: put one, get one
        inputs.put(align.TargetQueryPair(bases(c["target"]), bases(c["query"])))
        r = outputs.get(timeout=120)
        assert (r.Score, align.PrintCigar(r.Cigar)) == (c["score"], c["cigar"])
    for c in cases * 50:  # queued: results must come back in input order
        inputs.put(align.TargetQueryPair(bases(c["target"]), bases(c["query"])))
    for c in cases * 50:
        r = outputs.get(timeout=120)
        assert (r.Score, align.PrintCigar(r.Cigar)) == (c["score"], c["cigar"])
    inputs.close()
    assert outputs.get(timeout=120) is None


def test_engine_forwards_errors_and_closes():
    """A failing batch (a base >= dim: the reference goroutine would panic) must not leave consumers blocked: the
    error comes out of `outputs`, followed by the close sentinel."""
    inputs, outputs = align.GoAffineGapLocalEngine(MATRICES["Default"], -600, -150)
    inputs.put(align.TargetQueryPair(np.array([0, 1, 9, 3], dtype=np.uint8), np.array([0, 1, 2], dtype=np.uint8)))
    r = outputs.get(timeout=120)
    assert isinstance(r, _lib.GnxError) and r.code == _lib.GNX_EBASE
    assert outputs.get(timeout=120) is None


# ---- several GPUs behind one C call (gnx_multi_*) ---------------------------------------------
def _multi_devices(k):
    import ctypes as C
    n = _lib.load().gnx_device_count()
    return [d % max(n, 1) for d in range(k)]  # fewer GPUs than shards: several contexts share a device


@pytest.mark.parametrize("shards", [2, 3])
def test_multi_gpu_abi_matches_single_device(ctx, shards):
    """gnx_multi_affine_batch / gnx_multi_const_batch: contiguous cell-balanced shards on one context per device
    (the same device repeated when the box has fewer), results stitched in pair order == the oracle, for ragged
    batches, both modes, with and without cigars, tiny staging (GNX_ECAP inside a shard) and a too-small caller
    buffer (GNX_ECAP + gnx_multi_copy_last_cigars)."""
    rng = np.random.default_rng(77 + shards)
    al, be = [], []
    for k in range(997):
        n, m = int(rng.integers(0, 400)), int(rng.integers(0, 200))
        a, b = random_pair(rng, n, m, identity=float(rng.choice([0.7, 0.9, 1.0])))
        al.append(a)
        be.append(b)
    al.append(rng.integers(0, 4, 3000, dtype=np.uint8))  # one heavy pair at the end: unbalanced cuts
    be.append(rng.integers(0, 4, 2500, dtype=np.uint8))
    ac, ao = concat(al)
    bc, bo = concat(be)
    S = orc.HUMAN_CHIMP_TWO_SCORE_MATRIX
    with align.MultiContext(_multi_devices(shards)) as mc:
        assert mc.n_devices == shards
        cuts = mc.shard_bounds(ao, bo)
        from gonomics_b200 import shard as shard_mod
        assert [(int(cuts[r]), int(cuts[r + 1])) for r in range(shards)] == shard_mod.shard_bounds(ao, bo, shards)
        for mode in (0, 1, 2):
            O = -600 if mode != 2 else -430
            osc, ooff, ocig = orc.batch(ac, ao, bc, bo, S, O, -150, mode, True, 8)
            if mode == 2:
                sc, off, cig = mc.const_gap_batch(ac, ao, bc, bo, S, O, True)
                sc0, _, _ = mc.const_gap_batch(ac, ao, bc, bo, S, O, False)
            else:
                sc, off, cig = mc.affine_gap_batch(ac, ao, bc, bo, S, O, -150, mode == 1, True)
                sc0, _, _ = mc.affine_gap_batch(ac, ao, bc, bo, S, O, -150, mode == 1, False)
            assert np.array_equal(sc, osc) and np.array_equal(sc0, osc) and np.array_equal(off, ooff)
            assert np.array_equal(cig["run_length"], ocig["run_length"]) and np.array_equal(cig["op"], ocig["op"])
        # caller buffer too small: GNX_ECAP with scores and offsets filled, cigars fetched afterwards
        osc, ooff, ocig = orc.batch(ac, ao, bc, bo, S, -600, -150, 0, True, 8)
        out = (np.zeros(len(al), np.int64), np.zeros(len(al) + 1, np.int64), np.zeros(5, _lib.CIGAR_DTYPE))
        with pytest.raises(_lib.GnxError) as ei:
            mc.affine_gap_batch(ac, ao, bc, bo, S, -600, -150, False, True, out=out)
        assert ei.value.code == _lib.GNX_ECAP and np.array_equal(out[0], osc) and np.array_equal(out[1], ooff)
        big = np.zeros(int(out[1][-1]), dtype=_lib.CIGAR_DTYPE)
        assert mc._L.gnx_multi_copy_last_cigars(mc._h, big.ctypes.data, len(big)) == 0
        assert np.array_equal(big["run_length"], ocig["run_length"]) and np.array_equal(big["op"], ocig["op"])
        # an invalid base in the second shard is reported like the single-device call reports it
        bad = ac.copy()
        bad[ao[len(al) - 2] + 1] = 9
        with pytest.raises(_lib.GnxError) as ei:
            mc.affine_gap_batch(bad, ao, bc, bo, S, -600, -150, False, True)
        assert ei.value.code == _lib.GNX_EBASE
        # empty batch
        sc, off, cig = mc.affine_gap_batch(np.zeros(0, np.uint8), np.zeros(1, np.int64), np.zeros(0, np.uint8),
                                           np.zeros(1, np.int64), S, -600, -150, False, True)
        assert len(sc) == 0 and list(off) == [0]


# ---- dnaTwoBit inputs (gnx_affine_batch_twobit) -------------------------------------------------------
def _pack_ragged(seqs):
    """Tightly packed NewTwoBit words of every sequence (oracle packer = dnaTwoBit.NewTwoBit) + lengths."""
    words = [orc.new_twobit(s)[0][:(len(s) + 31) // 32] for s in seqs]
    cat = np.concatenate(words + [np.zeros(0, dtype=np.uint64)]).astype(np.uint64)
    return cat, np.array([len(s) for s in seqs], dtype=np.int64)


@pytest.mark.parametrize("tma", [1, 0])
@pytest.mark.parametrize("n,m", [(500, 150), (300, 141), (333, 142), (512, 160), (64, 33), (290, 145), (31, 150), (700, 150)])
def test_twobit_uniform_batches(n, m, tma):
    """Uniform 2-bit batches: the packed 16-bit kernels read the dnaTwoBit words staged by TMA (tma = 1) or a device
    expansion (tma = 0; also n > 512); score-only and checkpoint-path traceback, free-end and global, pair counts
    that leave a partial quad, several chunks -- all equal to the oracle on the unpacked bases."""
    from gonomics_b200.synth import pack_uniform
    c = align.Context(0)
    try:
        c.set_option("tb_tma", tma)
        c.set_option("chunk_pairs", 1000)
        S = orc.HUMAN_CHIMP_TWO_SCORE_MATRIX
        for P in (2503, 5):
            a, ao, b, bo = synth_pairs(300 + n + m, P, n, m)
            wa, wb = pack_uniform(a, P, n), pack_uniform(b, P, m)
            for mode in (1, 0):
                osc, ooff, ocig = orc.batch(a, ao, b, bo, S, -600, -150, mode, True, 8)
                sc, off, cig = c.affine_gap_batch_twobit(wa, n, wb, m, S, -600, -150, mode == 1, True, n_pairs=P)
                sc0, _, _ = c.affine_gap_batch_twobit(wa, n, wb, m, S, -600, -150, mode == 1, False, n_pairs=P)
                assert np.array_equal(sc, osc) and np.array_equal(sc0, osc), (n, m, P, mode)
                assert np.array_equal(off, ooff) and np.array_equal(cig["run_length"], ocig["run_length"]) \
                    and np.array_equal(cig["op"], ocig["op"]), (n, m, P, mode)
    finally:
        c.close()


@pytest.mark.parametrize("n,m", [(500, 150), (333, 142), (96, 64), (31, 150), (512, 160)])
def test_pageable_bytes_are_packed_while_staged(n, m):
    """gnx_affine_batch on a large uniform batch in pageable memory: the staging pass packs the bytes to dnaTwoBit
    words (pack_stage = 1, the default) and the chunk runs on the TMA-fed kernels; same scores and cigars as with the
    byte staging (pack_stage = 0) and as the oracle.  A base >= 4 anywhere makes the call fall back to the byte path,
    which reports the pair (GNX_EBASE); a 5 x 5 matrix (N is a legal base) never packs."""
    c = align.Context(0)
    try:
        c.set_option("chunk_pairs", 3000)
        S = orc.HUMAN_CHIMP_TWO_SCORE_MATRIX
        P = 8195  # several chunks, a partial quad at the end
        a, ao, b, bo = synth_pairs(900 + n + m, P, n, m)
        for mode in (1, 0):
            osc, ooff, ocig = orc.batch(a, ao, b, bo, S, -600, -150, mode, True, 8)
            for pack in (1, 0):
                c.set_option("pack_stage", pack)
                sc, off, cig = c.affine_gap_batch(a, ao, b, bo, S, -600, -150, mode == 1, True)
                sc0, _, _ = c.affine_gap_batch(a, ao, b, bo, S, -600, -150, mode == 1, False)
                assert np.array_equal(sc, osc) and np.array_equal(sc0, osc), (n, m, mode, pack)
                assert np.array_equal(off, ooff) and np.array_equal(cig["run_length"], ocig["run_length"]) \
                    and np.array_equal(cig["op"], ocig["op"]), (n, m, mode, pack)
        c.set_option("pack_stage", 1)
        bad = a.copy()
        bad[(P - 7) * n + n // 2] = 7  # in the last chunk: earlier chunks are already in flight when it is met
        with pytest.raises(_lib.GnxError) as ei:
            c.affine_gap_batch(bad, ao, b, bo, S, -600, -150, True, True)
        assert ei.value.code == _lib.GNX_EBASE
        # the context is still usable; an N (legal under the 5 x 5 matrix) sends the call to the byte path as well
        c.set_option("pack_stage", 1)  # (clears the back-off of the failed attempt)
        sc, _, _ = c.affine_gap_batch(a, ao, b, bo, S, -600, -150, True, False)
        assert np.array_equal(sc, orc.batch(a, ao, b, bo, S, -600, -150, 1, False, 8)[0])
        withn = a.copy()
        withn[(P - 7) * n + n // 2] = 4
        withn[5] = 4
        scn, offn, cign = c.affine_gap_batch(withn, ao, b, bo, S, -600, -150, True, True)
        on = orc.batch(withn, ao, b, bo, S, -600, -150, 1, True, 8)
        assert np.array_equal(scn, on[0]) and np.array_equal(offn, on[1]) and np.array_equal(cign["op"], on[2]["op"]) \
            and np.array_equal(cign["run_length"], on[2]["run_length"])
    finally:
        c.close()


def test_twobit_ragged_and_long(ctx):
    """Ragged 2-bit batches (per-pair lengths; empty sequences; multi-strip pairs) go through the device expansion and
    the ordinary kernels."""
    rng = np.random.default_rng(321)
    al, be = [], []
    for k in range(400):
        n, m = int(rng.integers(0, 300)), int(rng.integers(0, 200))
        a, b = random_pair(rng, n, m, identity=0.9)
        al.append(a)
        be.append(b)
    for n, m in [(3000, 1000), (700, 2000), (33, 0), (0, 65)]:
        a, b = random_pair(rng, n, m, identity=0.9)
        al.append(a)
        be.append(b)
    wa, la = _pack_ragged(al)
    wb, lb = _pack_ragged(be)
    ac, ao = concat(al)
    bc, bo = concat(be)
    S = orc.HUMAN_CHIMP_TWO_SCORE_MATRIX
    for mode in (0, 1):
        osc, ooff, ocig = orc.batch(ac, ao, bc, bo, S, -600, -150, mode, True, 8)
        sc, off, cig = ctx.affine_gap_batch_twobit(wa, la, wb, lb, S, -600, -150, mode == 1, True)
        assert np.array_equal(sc, osc) and np.array_equal(off, ooff)
        assert np.array_equal(cig["run_length"], ocig["run_length"]) and np.array_equal(cig["op"], ocig["op"])
        sc0, _, _ = ctx.affine_gap_batch_twobit(wa, la, wb, lb, S, -600, -150, mode == 1, False)
        assert np.array_equal(sc0, osc)


# ---- randomised differential tests -----------------------------------------------------------
PENALTIES = [(-400, -30), (-600, -150), (-300, -40), (-200, -50)]


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_random_short_ragged(ctx, mode):
    rng = np.random.default_rng(100 + mode)
    for mi, (name, S) in enumerate(MATRICES.items()):
        O, E = PENALTIES[mi % 4]
        al, be = [], []
        for _ in range(600):
            n, m = int(rng.integers(0, 90)), int(rng.integers(0, 90))
            a, b = random_pair(rng, n, m, identity=float(rng.uniform(0.5, 1.0)))
            al.append(a)
            be.append(b)
        check_batch(ctx, al, be, S, O if mode != 2 else -430, E, mode)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_random_medium_lengths(ctx, mode):
    rng = np.random.default_rng(200 + mode)
    S = orc.HUMAN_CHIMP_TWO_SCORE_MATRIX
    al, be = [], []
    for _ in range(300):
        n, m = int(rng.integers(1, 600)), int(rng.integers(1, 600))
        a, b = random_pair(rng, n, m, identity=float(rng.uniform(0.6, 1.0)))
        al.append(a)
        be.append(b)
    for n, m in [(1, 1), (1, 400), (400, 1), (160, 160), (161, 161), (500, 150), (150, 500), (320, 320), (321, 321),
                 (31, 5), (5, 31), (1000, 150), (33, 640), (640, 33)]:
        a, b = random_pair(rng, n, m, identity=0.9)
        al.append(a)
        be.append(b)
    check_batch(ctx, al, be, S, -600 if mode != 2 else -430, -150, mode)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_tie_heavy_inputs(ctx, mode):
    """Homopolymers and short tandem repeats: many equal-score paths, so any deviation from the
    M >= I >= D tie-break shows up in the cigar."""
    rng = np.random.default_rng(300 + mode)
    al, be = [], []
    for _ in range(400):
        unit = rng.integers(0, 4, int(rng.integers(1, 4)), dtype=np.uint8)
        n, m = int(rng.integers(1, 200)), int(rng.integers(1, 200))
        a = np.resize(unit, n).copy()
        b = np.resize(unit, m).copy()
        for arr in (a, b):  # a few point changes
            k = int(rng.integers(0, 4))
            if len(arr) and k:
                arr[rng.integers(0, len(arr), k)] = rng.integers(0, 4, k)
        al.append(a)
        be.append(b)
    for name, S in MATRICES.items():
        for O, E in ((-400, -30), (-100, -100), (0, -50), (-91, -91)):
            check_batch(ctx, al, be, S, O if mode != 2 else E, E, mode)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_pairs_with_N(ctx, mode):
    rng = np.random.default_rng(400 + mode)
    al, be = [], []
    for k in range(300):
        n, m = int(rng.integers(1, 300)), int(rng.integers(1, 300))
        a, b = random_pair(rng, n, m, identity=0.85)
        if k % 3:
            a[rng.integers(0, n, max(1, n // 20))] = 4
        if k % 2:
            b[rng.integers(0, m, max(1, m // 20))] = 4
        al.append(a)
        be.append(b)
    for name in ("Default", "HumanChimpTwo", "HoxD55"):
        check_batch(ctx, al, be, MATRICES[name], -600 if mode != 2 else -430, -150, mode)


def test_score_only_matches(ctx):
    rng = np.random.default_rng(500)
    al, be = [], []
    for _ in range(500):
        n, m = int(rng.integers(0, 400)), int(rng.integers(0, 400))
        a, b = random_pair(rng, n, m, identity=0.9)
        al.append(a)
        be.append(b)
    for mode in (0, 1, 2):
        check_batch(ctx, al, be, orc.HUMAN_CHIMP_TWO_SCORE_MATRIX, -600 if mode != 2 else -430, -150, mode,
                    want_cigar=False)


def test_long_pairs_multi_strip(ctx):
    rng = np.random.default_rng(600)
    al, be = [], []
    for n, m in [(3000, 2500), (2500, 3000), (1200, 5000), (5000, 700)]:
        a, b = random_pair(rng, n, m, identity=0.92)
        al.append(a)
        be.append(b)
    for mode in (0, 1, 2):
        check_batch(ctx, al, be, orc.HUMAN_CHIMP_TWO_SCORE_MATRIX, -600 if mode != 2 else -430, -150, mode)


@pytest.mark.parametrize("wide", [0, 1])
def test_long_pairs_cta_per_pair_kernel(wide):
    """The 4-warp CTA-per-pair kernel (affine_fill3w_kernel: strips pipelined over warps through shared-memory
    rings) and the one-warp multi-strip kernel give the oracle's result on ragged long pairs: 1..7 strips, more
    pairs than resident CTAs is not needed (each CTA loops), N bases, empty sides, both modes, with/without trace."""
    c = align.Context(0)
    try:
        c.set_option("wide_cta", wide)
        rng = np.random.default_rng(601 + wide)
        al, be = [], []
        shapes = [(1500, 100), (1030, 321), (40, 700), (900, 640), (2000, 1281), (333, 1600), (2100, 2100), (5, 330),
                  (1, 961), (1300, 0), (0, 1300), (64, 1990), (3000, 1000)]
        for rep in range(3):
            for n, m in shapes:
                a, b = random_pair(rng, n, m, identity=0.9)
                if rep == 1 and n and m:
                    a[rng.integers(0, n, max(1, n // 50))] = 4
                al.append(a)
                be.append(b)
        for mode in (0, 1):
            check_batch(c, al, be, orc.HUMAN_CHIMP_TWO_SCORE_MATRIX, -600, -150, mode)
            check_batch(c, al, be, orc.DEFAULT_SCORE_MATRIX, -400, -30, mode, want_cigar=False)
            check_batch(c, al, be, orc.DEFAULT_SCORE_MATRIX, 30, -40, mode, want_cigar=False)  # O > 0: tagged score-only
    finally:
        c.close()


@pytest.mark.parametrize("form", [0, 1])
def test_long_pairs_tile_checkpoint_kernel(form):
    """affine_long_kernel (gnx_long.cuh: score-only sweep with strip-edge columns + row checkpoints, then recompute of
    the route's tiles) forced onto ragged pairs: 1..16 strips, several row-checkpoint blocks, pairs shorter than one
    block, N bases, empty sides, both modes, both cell formulations of the score-only pass, cigars beyond the
    1024-entry slot (second pass) and a tie-heavy matrix."""
    c = align.Context(0)
    try:
        c.set_option("long_ckpt", 1)
        c.set_option("long_form", form)
        rng = np.random.default_rng(9100 + form)
        al, be = [], []
        shapes = [(1500, 100), (1030, 321), (40, 700), (900, 640), (2000, 1281), (333, 1600), (2100, 2100), (5, 330),
                  (1, 961), (1300, 0), (0, 1300), (64, 1990), (3000, 1000), (257, 321), (256, 640), (255, 330),
                  (513, 5000), (5000, 513), (226, 960), (289, 322)]
        for rep in range(3):
            for n, m in shapes:
                a, b = random_pair(rng, n, m, identity=float(rng.choice([0.6, 0.9, 0.97])))
                if rep == 1 and n and m:
                    a[rng.integers(0, n, max(1, n // 50))] = 4
                al.append(a)
                be.append(b)
        for mode in (0, 1):
            check_batch(c, al, be, orc.HUMAN_CHIMP_TWO_SCORE_MATRIX, -600, -150, mode)
        # cheap gaps on unrelated sequences: thousands of cigar elements per pair (slot overflow pass), ties everywhere
        al = [rng.integers(0, 4, 4000, dtype=np.uint8) for _ in range(6)] + [np.resize(np.array([0, 1], np.uint8), 3000)]
        be = [rng.integers(0, 4, 3900, dtype=np.uint8) for _ in range(6)] + [np.resize(np.array([1, 0, 0], np.uint8), 2800)]
        for mode in (0, 1):
            check_batch(c, al, be, orc.DEFAULT_SCORE_MATRIX, -20, -5, mode)
            check_batch(c, al, be, orc.DEFAULT_SCORE_MATRIX, 0, -30, mode)
        # a run pool too small for the batch: the pairs that find it full are re-run by the kernel's second pass
        c.set_option("long_pool", 3000)
        check_batch(c, al, be, orc.DEFAULT_SCORE_MATRIX, -20, -5, 0)
    finally:
        c.close()


@pytest.mark.parametrize("path", ["tile_checkpoint", "trace_matrix_warp", "trace_matrix_cta"])
def test_config_c4_10kb_pairs(path):
    """BASELINE configs[3] shape: 64 random 10 kb x 10 kb global pairs (the SURVEY 8d recipe: related with
    substitutions and indels, 10 % unrelated whose cigars run past the 1024-entry slot) + indel-heavy pairs + the
    reference's own 9673 x 10000 PanTro6/hg38 pair, bit-exact against the oracle on every long-pair path."""
    c = align.Context(0)
    try:
        c.set_option("long_ckpt", 1 if path == "tile_checkpoint" else 0)
        if path != "tile_checkpoint":
            c.set_option("wide_cta", 1 if path == "trace_matrix_cta" else 0)
        a, ao, b, bo = synth_pairs(20260104, 56, 10_000, 10_000)
        al = [a[ao[p]:ao[p + 1]] for p in range(56)]
        be = [b[bo[p]:bo[p + 1]] for p in range(56)]
        rng = np.random.default_rng(4242)
        for k in range(8):
            x, y = random_pair(rng, 10_000 - 7 * k, 10_000 - 13 * (k % 3), identity=0.7)
            al.append(x)
            be.append(y)
        g = load("cigar_to_bed")
        big = [cs for cs in g["cases"] if len(cs["alpha"]) > 9000][0]
        al.append(bases(big["alpha"], upper=True))
        be.append(bases(big["beta"], upper=True))
        sc = check_batch(c, al, be, orc.HUMAN_CHIMP_TWO_SCORE_MATRIX, -600, -150, 0, threads=16)
        assert int(sc[-1]) == 790738
    finally:
        c.close()


def _uniform_batch(rng, P, n, m, flavour):
    al, be = [], []
    for k in range(P):
        if flavour == "related":
            a, b = random_pair(rng, n, m, identity=float(rng.choice([0.6, 0.8, 0.9, 0.97, 1.0])))
        elif flavour == "ties":  # homopolymers / short tandem repeats: ties everywhere
            unit = rng.integers(0, 4, size=int(rng.integers(1, 4)), dtype=np.uint8)
            a = np.resize(unit, n).astype(np.uint8)
            b = np.resize(np.roll(unit, int(rng.integers(0, 3))), m).astype(np.uint8)
            if k % 3 == 0:
                b[rng.integers(0, m, 3)] = rng.integers(0, 4, 3)
        elif flavour == "indels":  # many short runs: cigars longer than the 24-entry slot
            a = rng.integers(0, 4, size=n, dtype=np.uint8)
            src = a[int(rng.integers(0, n - m)):]
            out, i = [], 0
            while len(out) < m:
                piece = src[i:i + 4].tolist()
                out.extend(piece if piece else rng.integers(0, 4, size=4).tolist())
                i += 4 + (1 if (len(out) // 4) % 2 else 0)
                if (len(out) // 4) % 3 == 0:
                    out.append(int(rng.integers(0, 4)))
            b = np.array(out[:m], dtype=np.uint8)
        else:  # unrelated
            a = rng.integers(0, 4, size=n, dtype=np.uint8)
            b = rng.integers(0, 4, size=m, dtype=np.uint8)
        if flavour == "related" and k % 7 == 3:
            a[rng.integers(0, n, 4)] = 4
            b[rng.integers(0, m, 2)] = 4
        al.append(a)
        be.append(b)
    return al, be


@pytest.mark.parametrize("n,m", [(500, 150), (300, 141), (333, 142), (400, 143), (1024, 144), (290, 145), (292, 146),
                                 (301, 147), (640, 148), (298, 149), (64, 30), (45, 9), (700, 160), (33, 1), (2, 1)])
def test_checkpoint_recompute_traceback(n, m):
    """Uniform freeEndGaps batches take the checkpoint-and-recompute path (fill16 + affine_ckpt_trace_kernel);
    it must give the oracle's score and cigar, and the same as the trace-matrix path (ckpt = 0)."""
    c = align.Context(0)
    try:
        rng = np.random.default_rng(7000 + n + m)
        for flavour, P, O, E, S in (("related", 203, -600, -150, orc.HUMAN_CHIMP_TWO_SCORE_MATRIX),
                                    ("ties", 101, -600, -150, orc.HUMAN_CHIMP_TWO_SCORE_MATRIX),
                                    ("ties", 50, 0, -30, orc.DEFAULT_SCORE_MATRIX),
                                    ("indels", 64, -100, -20, orc.DEFAULT_SCORE_MATRIX),
                                    ("unrelated", 64, -20, -5, orc.DEFAULT_SCORE_MATRIX),  # cheap gaps: > 64 ops
                                    ("unrelated", 97, -400, -30, orc.DEFAULT_SCORE_MATRIX)):
            if flavour == "indels" and n - m < 8:
                continue
            al, be = _uniform_batch(rng, P, n, m, flavour)
            if flavour == "unrelated" and O == -20 and m >= 140:  # these cigars must overflow the 64-entry slot
                ac, ao = concat(al)
                bc, bo = concat(be)
                _, ooff, _ = orc.batch(ac, ao, bc, bo, S, O, E, 1, True, 8)
                assert int(np.diff(ooff).max()) > 64
            c.set_option("ckpt", 1)
            sc = check_batch(c, al, be, S, O, E, 1)
            c.set_option("ckpt", 0)
            sc0 = check_batch(c, al, be, S, O, E, 1)
            assert np.array_equal(sc, sc0)
    finally:
        c.close()


def test_warp_per_pair_traceback_kernel():
    """traceback_affine_warp_kernel (a warp per pair, match runs consumed 32 cells at a time) is the default for
    long pairs; forced here (tb_impl = 3) onto every fill3 trace so that short, ragged, tie-heavy and N-bearing
    pairs, both modes, and cigars beyond the slot go through it too."""
    c = align.Context(0)
    try:
        c.set_option("tb_impl", 3)
        c.set_option("ckpt", 0)
        rng = np.random.default_rng(811)
        al, be = [], []
        for k in range(300):
            n, m = int(rng.integers(0, 700)), int(rng.integers(0, 700))
            a, b = random_pair(rng, n, m, identity=float(rng.choice([0.5, 0.8, 0.95, 1.0])))
            if k % 9 == 0 and n and m:
                a[rng.integers(0, n, 3)] = 4
            if k % 11 == 0:
                unit = rng.integers(0, 4, size=2, dtype=np.uint8)
                a, b = np.resize(unit, n).astype(np.uint8), np.resize(unit[::-1], m).astype(np.uint8)
            al.append(a)
            be.append(b)
        al += [rng.integers(0, 4, 3000, dtype=np.uint8), rng.integers(0, 4, 40, dtype=np.uint8)]
        be += [rng.integers(0, 4, 2500, dtype=np.uint8), rng.integers(0, 4, 1500, dtype=np.uint8)]
        for mode in (0, 1):
            check_batch(c, al, be, orc.HUMAN_CHIMP_TWO_SCORE_MATRIX, -600, -150, mode)
            check_batch(c, al, be, orc.DEFAULT_SCORE_MATRIX, -20, -5, mode)  # cheap gaps: dozens of ops per cigar
        a, ao, b, bo = synth_pairs(5, 2000, 500, 150)
        check_batch(c, [a[ao[p]:ao[p + 1]] for p in range(2000)], [b[bo[p]:bo[p + 1]] for p in range(2000)],
                    orc.HUMAN_CHIMP_TWO_SCORE_MATRIX, -600, -150, 1)
    finally:
        c.close()


def test_chunking_and_small_workspace():
    """Force many chunks (tiny workspace / chunk_pairs) so chunk boundaries and slot reuse are exercised."""
    c = align.Context(0, workspace_bytes=8 << 20)
    try:
        c.set_option("chunk_pairs", 37)
        a, ao, b, bo = synth_pairs(11, 1000, 200, 80)
        al = [a[ao[p]:ao[p + 1]] for p in range(1000)]
        be = [b[bo[p]:bo[p + 1]] for p in range(1000)]
        check_batch(c, al, be, orc.HUMAN_CHIMP_TWO_SCORE_MATRIX, -600, -150, 1)
        check_batch(c, al, be, orc.HUMAN_CHIMP_TWO_SCORE_MATRIX, -600, -150, 0, want_cigar=False)
    finally:
        c.close()


@pytest.mark.parametrize("n,m", [(120, 100), (300, 100)])  # trace-matrix path / checkpoint-and-recompute path
def test_cigar_cap_overflow_and_fetch(ctx, n, m):
    a, ao, b, bo = synth_pairs(12, 500, n, m)
    S = orc.HUMAN_CHIMP_TWO_SCORE_MATRIX
    want = ctx.affine_gap_batch(a, ao, b, bo, S, -600, -150, True)
    out_score = np.zeros(500, dtype=np.int64)
    out_off = np.zeros(501, dtype=np.int64)
    small = np.zeros(7, dtype=_lib.CIGAR_DTYPE)
    with pytest.raises(_lib.GnxError) as ei:
        ctx.affine_gap_batch(a, ao, b, bo, S, -600, -150, True, out=(out_score, out_off, small))
    assert ei.value.code == _lib.GNX_ECAP
    assert np.array_equal(out_score, want[0]) and np.array_equal(out_off, want[1])
    big = np.zeros(int(out_off[-1]), dtype=_lib.CIGAR_DTYPE)
    assert ctx._L.gnx_copy_last_cigars(ctx._h, big.ctypes.data, len(big)) == 0
    assert np.array_equal(big["run_length"], want[2]["run_length"]) and np.array_equal(big["op"], want[2]["op"])


def test_many_ops_overflow_slot(ctx):
    """Cigars longer than the per-pair slot take the second traceback pass."""
    rng = np.random.default_rng(700)
    al, be = [], []
    for _ in range(50):  # unrelated pairs with cheap gaps -> dozens of ops
        al.append(rng.integers(0, 4, 300, dtype=np.uint8))
        be.append(rng.integers(0, 4, 280, dtype=np.uint8))
    S = orc.DEFAULT_SCORE_MATRIX
    check_batch(ctx, al, be, S, -20, -5, 0)
    check_batch(ctx, al, be, S, -30, 0, 2)


def test_wide_int64_fallback(ctx):
    """Scores/penalties too large for the int32 range proof take the int64 instantiation: still bit-exact."""
    rng = np.random.default_rng(900)
    S = orc.HUMAN_CHIMP_TWO_SCORE_MATRIX * 1000
    al, be = [], []
    for _ in range(200):
        n, m = int(rng.integers(0, 400)), int(rng.integers(0, 400))
        a, b = random_pair(rng, n, m, identity=float(rng.uniform(0.6, 1.0)))
        if rng.random() < 0.3 and n:
            a[rng.integers(0, n)] = 4
        al.append(a)
        be.append(b)
    for mode in (0, 1):
        check_batch(ctx, al, be, S, -600_000, -150_000, mode)
        check_batch(ctx, al, be, S, -600_000, -150_000, mode, want_cigar=False)


def test_invalid_base_is_an_error(ctx):
    a = np.array([0, 1, 7, 3], dtype=np.uint8)  # LowerG: Go panics with index out of range
    b = np.array([0, 1, 2, 3], dtype=np.uint8)
    with pytest.raises(_lib.GnxError) as ei:
        align.AffineGap_highMem(a, b, orc.DEFAULT_SCORE_MATRIX, -400, -30, ctx)
    assert ei.value.code == _lib.GNX_EBASE
    # ... but not when the other side is empty (the matrix is never indexed)
    s, c = align.AffineGap_highMem(a, np.zeros(0, dtype=np.uint8), orc.DEFAULT_SCORE_MATRIX, -400, -30, ctx)
    assert (s, c) == (-400 - 4 * 30, [(4, 2)])


def test_empty_inputs(ctx):
    S = orc.DEFAULT_SCORE_MATRIX
    e = np.zeros(0, dtype=np.uint8)
    acg = bases("ACG")
    assert align.AffineGap_highMem(e, e, S, -400, -30, ctx) == (0, [(0, 0)])
    assert align.AffineGap_highMem(e, acg, S, -400, -30, ctx) == (-490, [(3, 1)])
    assert align.AffineGap_highMem(acg, e, S, -400, -30, ctx) == (-490, [(3, 2)])
    assert align.AffineGapLocal(acg, e, S, -400, -30, ctx) == (0, [(3, 2)])
    assert align.ConstGap_highMem(e, e, S, -430, ctx) == (0, [(0, 0)])
    assert align.ConstGap_highMem(e, acg, S, -430, ctx) == (-1290, [(3, 1)])
    with pytest.raises(_lib.GnxError):
        align.AffineGap(e, acg, S, -400, -30, ctx)
    sc, off, cig = ctx.affine_gap_batch(e, np.zeros(1, dtype=np.int64), e, np.zeros(1, dtype=np.int64), S, -400, -30)
    assert len(sc) == 0 and list(off) == [0]


def test_device_resident_entry_point(ctx):
    import torch
    a, ao, b, bo = synth_pairs(13, 4000, 500, 150)
    S = orc.HUMAN_CHIMP_TWO_SCORE_MATRIX
    dev = torch.device("cuda:0")
    ta, tb = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
    tao, tbo = torch.from_numpy(ao).to(dev), torch.from_numpy(bo).to(dev)
    for kind, want_cigar in ((1, True), (1, False), (0, True), (2, True)):
        score = torch.zeros(4000, dtype=torch.int64, device=dev)
        off = torch.zeros(4001, dtype=torch.int64, device=dev)
        cap = 4000 * 700  # linear-gap global alignments of 500 vs 150 scatter their 350 deletions
        cig = torch.zeros(cap * 16, dtype=torch.uint8, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        stream = torch.cuda.current_stream().cuda_stream
        ctx.batch_device(kind, ta.data_ptr(), tao.data_ptr(), tb.data_ptr(), tbo.data_ptr(), ao, bo, 4000, S, -600,
                         -150, want_cigar, score.data_ptr(), cig.data_ptr(), off.data_ptr(), cap,
                         status.data_ptr(), stream)
        torch.cuda.synchronize()
        assert int(status.item()) == 0
        osc, ooff, ocig = orc.batch(a, ao, b, bo, S, -600, -150, kind, want_cigar, 8)
        assert np.array_equal(score.cpu().numpy(), osc)
        if want_cigar:
            assert np.array_equal(off.cpu().numpy(), ooff)
            got = cig.cpu().numpy()[:int(ooff[-1]) * 16].view(_lib.CIGAR_DTYPE)
            assert np.array_equal(got["run_length"], ocig["run_length"]) and np.array_equal(got["op"], ocig["op"])


# ---- BASELINE-shaped workloads -----------------------------------------------------------------
def test_config_c2_c3_prefix_and_properties(ctx):
    """C2/C3 shape (target 500 x query 150, free end gaps): the first 10^4 pairs are diffed against the
    oracle; the whole 2*10^5-pair batch is checked through size-independent properties: every cigar
    consumes exactly n target and m query bases, score-only equals the traceback run's scores."""
    N = 200_000
    a, ao, b, bo = synth_pairs(20260103, N, 500, 150)
    S = orc.HUMAN_CHIMP_TWO_SCORE_MATRIX
    sc, off, cig = ctx.affine_gap_batch(a, ao, b, bo, S, -600, -150, True)
    sc2, _, _ = ctx.affine_gap_batch(a, ao, b, bo, S, -600, -150, True, want_cigar=False)
    assert np.array_equal(sc, sc2)
    K = 10_000
    osc, ooff, ocig = orc.batch(a[:K * 500], ao[:K + 1], b[:K * 150], bo[:K + 1], S, -600, -150, 1, True, 8)
    assert np.array_equal(sc[:K], osc)
    assert np.array_equal(off[:K + 1], ooff)
    assert np.array_equal(cig["run_length"][:ooff[-1]], ocig["run_length"])
    assert np.array_equal(cig["op"][:ooff[-1]], ocig["op"])
    pair_of = np.repeat(np.arange(N), np.diff(off))
    rl, op = cig["run_length"], cig["op"]
    assert np.array_equal(np.bincount(pair_of, weights=rl * (op != 1), minlength=N).astype(np.int64), np.full(N, 500))
    assert np.array_equal(np.bincount(pair_of, weights=rl * (op != 2), minlength=N).astype(np.int64), np.full(N, 150))
    assert np.all(rl > 0) and np.all(op <= 2)
    assert np.all((op[1:] != op[:-1]) | (pair_of[1:] != pair_of[:-1]))  # runs are maximal


def test_config_c1_global_1kb(ctx):
    a, ao, b, bo = synth_pairs(20260101, 1000, 1000, 150)
    al = [a[ao[p]:ao[p + 1]] for p in range(1000)]
    be = [b[bo[p]:bo[p + 1]] for p in range(1000)]
    check_batch(ctx, al, be, orc.HUMAN_CHIMP_TWO_SCORE_MATRIX, -600, -150, 0)


# ---- profile DP / progressive multiple alignment ---------------------------------
def _random_group(rng, nseq, length, gap_frac=0.15, lower_frac=0.1, with_n=True):
    g = rng.integers(0, 5 if with_n else 4, size=(nseq, length), dtype=np.uint8)
    low = rng.random((nseq, length)) < lower_frac
    g[low] += 5  # lowercase bases fold to uppercase in scoreColumnMatch
    gaps = rng.random((nseq, length)) < gap_frac
    g[gaps] = 10
    g[0, :][g[0, :] == 10] = 1  # keep one ungapped base per column: no division by zero
    return g


def test_multi_affine_golden_fixtures(ctx):  # align/multiAlign_test.go:17-37 TestMultiAlignGap
    g = load("multi_align")
    S = MATRICES[g["matrix"]]
    for c in g["cases"]:
        recs = [(n, bases(s)) for n, s in c["input"]]
        want = sorted((n, s) for n, s in c["expected"])
        for chunk in (1, g["chunk"]):
            got = align.AllSeqAffineChunk(recs, S, g["gap_open"], g["gap_extend"], chunk, ctx)
            assert sorted((f.Name, orc.bases_to_string(f.Seq)) for f in got) == want, chunk
    g = load("affine_global")  # TestAffineGapMulti (affineGap_test.go:95-108): groups of one sequence
    S = MATRICES[g["matrix"]]
    for c in g["cases"]:
        a, b = bases(c["alpha"]), bases(c["beta"])
        sc, cig = align.multipleAffineGap([align.Fasta("one", a)], [align.Fasta("two", b)], S, g["gap_open"],
                                          g["gap_extend"], ctx)
        assert (sc, cig) == align.AffineGap_highMem(a, b, S, g["gap_open"], g["gap_extend"], ctx)
        assert align.View(a, b, cig) == c["view"]


def test_multi_affine_random_groups(ctx):
    rng = np.random.default_rng(815)
    for chunk in (1, 2, 3, 5):
        groups = []
        for _ in range(9):
            L = int(rng.integers(0, 70)) * chunk
            groups.append(_random_group(rng, int(rng.integers(1, 7)), L))
        groups.append(_random_group(rng, 3, 230 * chunk))  # more than one 160-column strip
        groups.append(_random_group(rng, 2, 400 * chunk))
        xs = [x for x in range(len(groups) - 1) for _ in range(x + 1, len(groups))]
        ys = [y for x in range(len(groups) - 1) for y in range(x + 1, len(groups))]
        for S, O, E in ((orc.DEFAULT_SCORE_MATRIX, -400, -30), (orc.HUMAN_CHIMP_TWO_SCORE_MATRIX, -600, -150)):
            sc, off, cig = ctx.multi_affine_chunk_batch(groups, xs, ys, S, O, E, chunk)
            sc2, _, _ = ctx.multi_affine_chunk_batch(groups, xs, ys, S, O, E, chunk, want_cigar=False)
            assert np.array_equal(sc, sc2)
            for p, (x, y) in enumerate(zip(xs, ys)):
                want = orc.multi_affine_gap_chunk(groups[x], groups[y], S, O, E, chunk)
                got = (int(sc[p]), [(int(r), int(o)) for r, o in cig[off[p]:off[p + 1]]])
                assert got == want, (chunk, x, y, groups[x].shape, groups[y].shape)


def test_multi_affine_small_workspace_and_errors(ctx):
    rng = np.random.default_rng(816)
    groups = [_random_group(rng, 3, 120) for _ in range(8)]
    xs = [x for x in range(7) for _ in range(x + 1, 8)]
    ys = [y for x in range(7) for y in range(x + 1, 8)]
    S = orc.DEFAULT_SCORE_MATRIX
    want = [orc.multi_affine_gap_chunk(groups[x], groups[y], S, -400, -30, 1) for x, y in zip(xs, ys)]
    c2 = align.Context(0, workspace_bytes=1 << 20)  # forces many sub-batches
    try:
        sc, off, cig = c2.multi_affine_chunk_batch(groups, xs, ys, S, -400, -30, 1, cigar_cap=8)  # + GNX_ECAP path
        got = [(int(sc[p]), [(int(r), int(o)) for r, o in cig[off[p]:off[p + 1]]]) for p in range(len(xs))]
        assert got == want
    finally:
        c2.close()
    # all-gap column pair: Go divides by zero
    bad = [g.copy() for g in groups]
    bad[2][:, 7] = 10
    bad[5][:, 3] = 10
    with pytest.raises(_lib.GnxError) as ei:
        ctx.multi_affine_chunk_batch(bad, xs, ys, S, -400, -30, 1)
    assert ei.value.code == _lib.GNX_EDIVZERO
    # a base outside the matrix (dna.Dot = 11) opposite an ungapped base: index out of range
    bad = [g.copy() for g in groups]
    bad[1][1, 5] = 11
    with pytest.raises(_lib.GnxError) as ei:
        ctx.multi_affine_chunk_batch(bad, xs, ys, S, -400, -30, 1)
    assert ei.value.code == _lib.GNX_EBASE
    with pytest.raises(orc.OracleError):
        orc.multi_affine_gap_chunk(bad[1], bad[0], S, -400, -30, 1)
    with pytest.raises(_lib.GnxError) as ei:  # length not a multiple of chunkSize: log.Fatalf
        ctx.multi_affine_chunk_batch(groups, xs, ys, S, -400, -30, 7)
    assert ei.value.code == _lib.GNX_ECHUNK


def test_all_seq_affine_random_families(ctx):
    """Whole progressive alignments (AllSeqAffine / AllSeqAffineChunk) against the oracle's driver."""
    from oracle import msa
    rng = np.random.default_rng(817)
    S = orc.DEFAULT_SCORE_MATRIX
    for chunk in (1, 4):
        root = rng.integers(0, 4, size=30 * chunk, dtype=np.uint8)
        recs = []
        for k in range(6):
            s = root.copy()
            for _ in range(int(rng.integers(0, 3))):  # delete / duplicate whole chunk-sized units
                u = int(rng.integers(0, len(s) // chunk)) * chunk
                s = np.delete(s, np.s_[u:u + chunk]) if rng.random() < 0.5 else np.insert(s, u, s[u:u + chunk])
            sub = rng.random(len(s)) < 0.05
            s[sub] = rng.integers(0, 4, size=int(sub.sum()), dtype=np.uint8)
            recs.append((f"s{k}", s))
        got = align.AllSeqAffineChunk(recs, S, -400, -30, chunk, ctx)
        want = msa.all_seq_affine_chunk(recs, S, -400, -30, chunk)
        assert [(f.Name, f.Seq.tolist()) for f in got] == [(n, s.tolist()) for n, s in want]


# ---- gsw extend step: LeftDynamicAln / RightDynamicAln ----------------------------
def _check_extend(ctx, alphas, betas, S, g):
    from gonomics_b200 import genomegraph as gg
    for side, ofn in ((1, orc.left_dynamic_aln), (2, orc.right_dynamic_aln)):
        got = gg.extend_pairs(side, alphas, betas, S, g, ctx)
        ac, ao = concat(alphas)
        bc, bo = concat(betas)
        sc_only = ctx.extend_batch(side, ac, ao, bc, bo, S, g, want_cigar=False)
        for p, (a, b) in enumerate(zip(alphas, betas)):
            want = ofn(a, b, S, g)
            assert (got[p][0], [tuple(c) for c in got[p][1]], got[p][2], got[p][3]) == want, \
                (side, p, len(a), len(b), g, got[p], want)
            assert int(sc_only[0][p]) == want[0]
            if side == 2:
                assert (int(sc_only[1][p]), int(sc_only[2][p])) == (want[2], want[3])


def test_extend_random_read_sized(ctx):
    rng = np.random.default_rng(2340)
    S = orc.HUMAN_CHIMP_TWO_SCORE_MATRIX
    alphas, betas = [], []
    for _ in range(3000):  # the gsw shapes: target window <= ~175, read remainder <= 150
        n, m = int(rng.integers(0, 180)), int(rng.integers(0, 151))
        a, b = random_pair(rng, max(n, 1), max(m, 1), identity=float(rng.choice([0.7, 0.9, 0.97, 1.0])))
        alphas.append(a[:n])
        betas.append(b[:m])
    _check_extend(ctx, alphas, betas, S, -600)
    _check_extend(ctx, alphas[:500], betas[:500], S, -100)
    _check_extend(ctx, alphas[:500], betas[:500], orc.DEFAULT_SCORE_MATRIX, 0)


def test_extend_ties_n_and_long(ctx):
    rng = np.random.default_rng(2341)
    S = orc.HUMAN_CHIMP_TWO_SCORE_MATRIX
    alphas, betas = [], []
    for unit in ("A", "AC", "ACG", "AAC"):  # tie-heavy repeats
        for n, m in ((30, 30), (64, 17), (17, 64), (150, 150)):
            alphas.append(bases((unit * 200)[:n]))
            betas.append(bases((unit * 200)[1:m + 1]))
    for _ in range(40):  # N bases (dim 5)
        a, b = random_pair(rng, int(rng.integers(1, 170)), int(rng.integers(1, 150)), alphabet=5)
        alphas.append(a)
        betas.append(b)
    _check_extend(ctx, alphas, betas, S, -600)
    _check_extend(ctx, alphas, betas, orc.DEFAULT_SCORE_MATRIX, -430)
    # more than one 160- / 320-column strip, and targets longer than the shared-memory stage
    alphas, betas = [], []
    for n, m in ((400, 170), (170, 400), (1500, 700), (700, 1500), (2000, 330)):
        a, b = random_pair(rng, n, m, identity=0.92)
        alphas.append(a)
        betas.append(b)
    _check_extend(ctx, alphas, betas, S, -600)
    with pytest.raises(_lib.GnxError) as ei:  # a base outside the matrix: Go panics
        _check_extend(ctx, [np.array([0, 1, 7], np.uint8)], [np.array([0, 1], np.uint8)], S, -600)
    assert ei.value.code == _lib.GNX_EBASE


def test_left_right_local_older_forms(ctx):
    """genomeGraph.LeftLocal / RightLocal (localAlignment.go:95-196) as modes of gnx_extend_batch: '=' / 'X' ops and the
    route in alignment order, against the plain-Python restatement (oracle/local.py)."""
    from oracle import local as oloc
    from gonomics_b200 import genomegraph
    from gonomics_b200._lib import GNX_EXT_LEFT_LOCAL, GNX_EXT_RIGHT_LOCAL
    rng = np.random.default_rng(4321)
    S = orc.HUMAN_CHIMP_TWO_SCORE_MATRIX
    al, be = [], []
    for t in range(160):
        n, m = int(rng.integers(0, 60)), int(rng.integers(0, 50))
        a, b = random_pair(rng, n, m, identity=float(rng.choice([0.6, 0.9, 1.0])))
        if t % 7 == 0:
            unit = rng.integers(0, 4, size=2, dtype=np.uint8)
            a, b = np.resize(unit, n).astype(np.uint8), np.resize(unit[::-1], m).astype(np.uint8)
        if t % 9 == 0 and n:
            a[int(rng.integers(0, n))] = 4
        al.append(a)
        be.append(b)
    for side, fn in ((GNX_EXT_LEFT_LOCAL, oloc.left_local), (GNX_EXT_RIGHT_LOCAL, oloc.right_local)):
        got = genomegraph.extend_pairs(side, al, be, S, -600, ctx)
        for p, (a, b) in enumerate(zip(al, be)):
            sc, route, min_i, max_i, min_j, max_j = fn(a, b, S, -600)
            end = (min_i, min_j) if side == GNX_EXT_LEFT_LOCAL else (max_i, max_j)
            assert (got[p][0], [tuple(x) for x in got[p][1]], got[p][2], got[p][3]) == (sc, route, end[0], end[1]), (side, p, len(a), len(b))
    a, b = al[3], be[3]
    assert genomegraph.LeftLocal(a, b, S, -600, ctx=ctx)[0] == oloc.left_local(a, b, S, -600)[0]
    assert genomegraph.RightLocal(a, b, S, -600, ctx=ctx)[3] == oloc.right_local(a, b, S, -600)[3]


def _ragged_reads(rng, P, nlo, nhi, mlo, mhi):
    al, be = [], []
    for _ in range(P):
        n, m = int(rng.integers(nlo, nhi + 1)), int(rng.integers(mlo, mhi + 1))
        a, b = random_pair(rng, n, m, identity=float(rng.choice([0.75, 0.95, 1.0])))
        al.append(a)
        be.append(b)
    return al, be


def test_ragged_batches_on_the_packed_16bit_kernels():
    """Ragged read batches (per-pair n and m) run on affine_fill16 (score only) and on the checkpoint-and-recompute
    path (traceback): quads binned on the host by (last-column index, target length).  Checked against the oracle and
    by the kernel path the library reports; several chunks; 2-bit ragged input; the int32 kernels as the other side."""
    rng = np.random.default_rng(1717)
    al, be = _ragged_reads(rng, 3001, 300, 500, 100, 150)
    al += [al[0][:301].copy(), al[1][:300].copy()]   # bins with a single pair
    be += [be[0][:150].copy(), be[1][:101].copy()]
    S = orc.HUMAN_CHIMP_TWO_SCORE_MATRIX
    c = align.Context(0)
    try:
        c.set_option("chunk_pairs", 1100)
        check_batch(c, al, be, S, -600, -150, 1, want_cigar=True)
        assert c.last_kernel_path() == (17, 1)
        check_batch(c, al, be, S, -600, -150, 1, want_cigar=False)
        assert c.last_kernel_path() == (16, 1)
        # global mode, score only (the 16-bit range proof holds for short targets and cheap extensions)
        gl, gb = _ragged_reads(rng, 800, 20, 200, 5, 160)
        check_batch(c, gl, gb, orc.DEFAULT_SCORE_MATRIX, -400, -30, 0, want_cigar=False)
        assert c.last_kernel_path() == (16, 1)
        # the same batch with the ragged path switched off: the int32 kernels
        c.set_option("ragged16", 0)
        check_batch(c, al[:500], be[:500], S, -600, -150, 1, want_cigar=True)
        assert c.last_kernel_path()[0] == 3
        c.set_option("ragged16", 1)
        # ragged 2-bit input: device expansion, then the same binned quads
        wa, la = _pack_ragged(al)
        wb, lb = _pack_ragged(be)
        ac, ao = concat(al)
        bc, bo = concat(be)
        osc, ooff, ocig = orc.batch(ac, ao, bc, bo, S, -600, -150, 1, True, 8)
        sc, off, cig = c.affine_gap_batch_twobit(wa, la, wb, lb, S, -600, -150, True, True)
        assert c.last_kernel_path() == (17, 1)
        assert np.array_equal(sc, osc) and np.array_equal(off, ooff)
        assert np.array_equal(cig["run_length"], ocig["run_length"]) and np.array_equal(cig["op"], ocig["op"])
    finally:
        c.close()
