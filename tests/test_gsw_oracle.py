"""Pin what can be pinned of the gsw oracles without a Go toolchain (the reference has no asserting test for these
functions): two INDEPENDENT restatements of the extension DPs held against each other (C in oracle/gnx_oracle.c,
plain Python in oracle/local.py), the older LeftLocal / RightLocal forms tied to them, and closed-form properties of the
per-read driver restatement (oracle/gsw.py).  CPU only."""
import numpy as np

import oracle as orc
from oracle import gsw as ogsw
from oracle import local as oloc

MATS = [orc.HUMAN_CHIMP_TWO_SCORE_MATRIX, orc.DEFAULT_SCORE_MATRIX]


def _pairs(rng, count, nmax, mmax):
    for t in range(count):
        n, m = int(rng.integers(0, nmax)), int(rng.integers(0, mmax))
        a = rng.integers(0, 4, size=n, dtype=np.uint8)
        if t % 3 == 0 and n and m:  # related pair
            b = np.resize(a, m).copy()
            for k in rng.integers(0, m, size=max(1, m // 8)):
                b[k] = (b[k] + 1) % 4
        elif t % 3 == 1:  # tie-heavy: short periodic sequences
            unit = rng.integers(0, 4, size=int(rng.integers(1, 4)), dtype=np.uint8)
            a, b = np.resize(unit, n).astype(np.uint8), np.resize(unit[::-1], m).astype(np.uint8)
        else:
            b = rng.integers(0, 4, size=m, dtype=np.uint8)
        if t % 10 == 9 and n:
            a[int(rng.integers(0, n))] = 4  # N
        yield a, b


def test_extension_dps_two_independent_restatements_agree():
    rng = np.random.default_rng(99)
    n_checked = 0
    for S in MATS:
        for g in (-600, -100, 0):
            for a, b in _pairs(rng, 120, 28, 24):
                assert orc.left_dynamic_aln(a, b, S, g) == tuple(oloc.left_dynamic_aln(a, b, S, g)), (a, b, g)
                assert orc.right_dynamic_aln(a, b, S, g) == tuple(oloc.right_dynamic_aln(a, b, S, g)), (a, b, g)
                n_checked += 1
    assert n_checked == 720


def test_local_forms_are_the_dynamic_forms_with_extended_ops():
    """LeftLocal / RightLocal = Left / RightDynamicAln with each M run split into '=' (positive substitution score) and 'X'
    runs and the route reversed; same score and end cells."""
    rng = np.random.default_rng(100)
    S = orc.HUMAN_CHIMP_TWO_SCORE_MATRIX
    for a, b in _pairs(rng, 150, 30, 26):
        for dyn, loc, left in ((oloc.left_dynamic_aln, oloc.left_local, True), (oloc.right_dynamic_aln, oloc.right_local, False)):
            sc, route, i, j = dyn(a, b, S, -600)
            lsc, lroute, min_i, max_i, min_j, max_j = loc(a, b, S, -600)
            assert lsc == sc
            assert (min_i, min_j) == (i, j) if left else (max_i, max_j) == (i, j)
            # rebuild the extended route from the dynamic one by walking it over the bases
            ci, cj = (len(a), len(b)) if left else (i, j)
            ops = []
            for run, op in route:
                for _ in range(run):
                    if op == "M":
                        ops.append("=" if S[a[ci - 1]][b[cj - 1]] > 0 else "X")
                        ci, cj = ci - 1, cj - 1
                    elif op == "I":
                        ops.append("I")
                        cj -= 1
                    else:
                        ops.append("D")
                        ci -= 1
            rle = []
            for o in ops[::-1]:
                if rle and rle[-1][1] == o:
                    rle[-1][0] += 1
                else:
                    rle.append([1, o])
            assert [tuple(x) for x in rle] == lroute


def test_gsw_driver_restatement_closed_forms():
    """Reads cut from the reference: a perfect read is one full-length seed (cigar 150M, score = perfectMatchBig, no
    extension); one substitution splits it into two seeds and the extension recovers the rest; the reverse strand maps
    with PosStrand false; junk maps nowhere; the pair flags follow setGirafFlags."""
    rng = np.random.default_rng(101)
    node = rng.integers(0, 4, size=20_000, dtype=np.uint8)
    gg = ogsw.LinearGenome([node], 32, 32)
    S = orc.HUMAN_CHIMP_TWO_SCORE_MATRIX
    for s in (0, 777, 12_345, 20_000 - 150):
        read = node[s:s + 150].copy()
        perfect = int(sum(S[b][b] for b in read))
        g = ogsw.graph_smith_waterman_to_giraf(gg, read, S)
        assert (g.AlnScore, g.Cigar, g.TStart, g.TEnd, g.PosStrand, g.QStart, g.Nodes, g.Flag) == \
            (perfect, ((150, "M"),), s, s + 150, True, 0, (0,), 4)
        r = ogsw.graph_smith_waterman_to_giraf(gg, orc.reverse_complement(read), S)
        assert (r.AlnScore, r.Cigar, r.TStart, r.PosStrand, r.Flag) == (perfect, ((150, "M"),), s, False, 0)
        if 200 < s < 19_000:
            sub = read.copy()
            sub[70] = (sub[70] + 1) % 4
            want = perfect - int(S[read[70]][read[70]]) + int(S[node[s + 70]][sub[70]])
            g = ogsw.graph_smith_waterman_to_giraf(gg, sub, S)
            assert (g.AlnScore, g.Cigar, g.TStart, g.TEnd) == (want, ((150, "M"),), s, s + 150)
            f, v = ogsw.wrap_pair_giraf(gg, read, orc.reverse_complement(node[s + 200:s + 350]), S)
            assert f.Flag == 4 + 8 + 16 + 16 + 1 and v.Flag == 0 + 1  # proper pair: Fwd before Rev, opposite strands
    junk = ogsw.graph_smith_waterman_to_giraf(gg, rng.integers(0, 4, size=150, dtype=np.uint8), S)
    assert junk.AlnScore < 1200 and junk.Flag & 2
    assert ogsw.seed_could_be_better(150, 0, 14_000, 150, 100, 90, -196, -296)
    assert not ogsw.seed_could_be_better(32, 14_000, 14_250, 150, 100, 90, -196, -296)
