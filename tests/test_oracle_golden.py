"""Pin the CPU oracle against every golden vector the reference holds for the hot path
(SURVEY.md 8c items 1-10).  CPU only."""
import numpy as np
import pytest

import oracle as orc
from oracle import msa
from golden_util import MATRICES, bases, cigar_to_beds, load, random_pair


def test_affine_global_views():  # align/affineGap_test.go:45-55 TestAffineGap
    g = load("affine_global")
    S = MATRICES[g["matrix"]]
    for c in g["cases"]:
        a, b = bases(c["alpha"]), bases(c["beta"])
        _, cig = orc.affine_gap_highmem(a, b, S, g["gap_open"], g["gap_extend"])
        assert orc.view(a, b, cig) == c["view"], c


def test_affine_lowmem_matches_highmem():  # align/affineGap_test.go:57-81 TestAffineGap_lowMem
    g = load("affine_global")
    S = MATRICES[g["matrix"]]
    for c in g["cases"]:
        a, b = bases(c["alpha"]), bases(c["beta"])
        hs, hc = orc.affine_gap_highmem(a, b, S, g["gap_open"], g["gap_extend"])
        ls, lc = orc.affine_gap_lowmem(a, b, S, g["gap_open"], g["gap_extend"])
        cs, cc = orc.affine_gap_lowmem(a, b, S, g["gap_open"], g["gap_extend"], 3, 3)
        assert ls == hs and cs == hs
        # the Go test compares element-wise over the low-mem route's length
        assert all(lc[i] == hc[i] for i in range(len(lc)))
        assert all(cc[i] == hc[i] for i in range(len(cc)))
        assert lc == hc  # single board: identical


def test_affine_chunk_views():  # align/affineGap_test.go:83-93 TestAffineGapChunk
    g = load("affine_chunk")
    S = MATRICES[g["matrix"]]
    for c in g["cases"]:
        a, b = bases(c["alpha"]), bases(c["beta"])
        _, cig = orc.affine_gap_chunk(a, b, S, g["gap_open"], g["gap_extend"], g["chunk"])
        assert orc.view(a, b, cig) == c["view"], c


def test_affine_multi_pairs():  # align/affineGap_test.go:95-108 TestAffineGapMulti
    g = load("affine_global")
    S = MATRICES[g["matrix"]]
    for c in g["cases"]:
        a, b = bases(c["alpha"]), bases(c["beta"])
        _, cig = orc.multi_affine_gap_chunk(a[None, :], b[None, :], S, g["gap_open"], g["gap_extend"], 1)
        merged = msa.merge_multiple_alignments([("one", a)], [("two", b)], cig)
        pretty = orc.bases_to_string(merged[0][1]) + "\n" + orc.bases_to_string(merged[1][1]) + "\n"
        assert pretty == c["view"], c


def test_affine_local_score_and_cigar():  # align/affineGap_test.go:120-155 TestAffineGapLocal
    for c in load("affine_local")["cases"]:
        score, cig = orc.affine_gap_local(bases(c["target"]), bases(c["query"]), MATRICES[c["matrix"]],
                                          c["gap_open"], c["gap_extend"])
        assert score == c["score"] and orc.print_cigar(cig) == c["cigar"], c


def test_const_gap_views():  # align/view_test.go:26-38 TestConstGap (low-mem driver) + high-mem
    g = load("const_gap")
    S = MATRICES[g["matrix"]]
    for c in g["cases"]:
        a, b = bases(c["alpha"]), bases(c["beta"])
        s1, cig = orc.const_gap_lowmem(a, b, S, g["gap_pen"])
        assert orc.view(a, b, cig) == c["view"], c
        s2, cig2 = orc.const_gap_highmem(a, b, S, g["gap_pen"])
        assert (s1, cig) == (s2, cig2)
        s3, cig3 = orc.const_gap_lowmem(a, b, S, g["gap_pen"], 3, 3)
        assert s3 == s2


def test_global_alignment_cmd():  # cmd/globalAlignment: ConstGap -> 3M3D3M
    g = load("global_alignment")
    a, b = bases(g["alpha"]), bases(g["beta"])
    score, cig = orc.const_gap_lowmem(a, b, MATRICES[g["matrix"]], g["gap_pen"])
    assert orc.view(a, b, cig) == g["view"]
    assert orc.print_cigar(cig) == "3M3D3M" and score == -730


def test_anchor_score_and_cigar():  # cmd/globalAlignmentAnchor out_alignment.{1,2}.expected.tsv
    g = load("anchor")
    S = MATRICES[g["matrix"]]
    assert len(g["cases"]) == 5
    for c in g["cases"]:
        a, b = bases(c["alpha"], upper=True), bases(c["beta"], upper=True)
        score, cig = orc.affine_gap_lowmem(a, b, S, g["gap_open"], g["gap_extend"], 10000, 10000)
        assert score == c["score"], c["region1"]
        assert [list(x) for x in cig] == c["cigar"], c["region1"]
        assert orc.affine_gap_highmem(a, b, S, g["gap_open"], g["gap_extend"]) == (score, cig)


def test_cigar_to_bed():  # cmd/cigarToBed TestCigarToBed (9 x 15 and 9673 x 10000)
    g = load("cigar_to_bed")
    S = MATRICES[g["matrix"]]
    for c in g["cases"]:
        a, b = bases(c["alpha"], upper=True), bases(c["beta"], upper=True)
        score, cig = orc.affine_gap_lowmem(a, b, S, g["gap_open"], g["gap_extend"])
        ins, dele = cigar_to_beds(cig, c["first_pos_ins"], c["first_pos_del"], c["chrom"])
        assert ins == c["ins_bed"] and dele == c["del_bed"], c["files"]
        if len(a) > 9000:
            assert (len(a), len(b)) == (9673, 10000)
            assert score == 790738 and len(cig) == 19
            assert orc.affine_gap_highmem(a, b, S, g["gap_open"], g["gap_extend"]) == (score, cig)
        else:
            assert score == -1070 and orc.print_cigar(cig) == "5M6I4M"


def test_multi_align_fixtures():  # align/multiAlign_test.go:17-37 TestMultiAlignGap
    g = load("multi_align")
    S = MATRICES[g["matrix"]]
    for c in g["cases"]:
        recs = [(n, bases(s)) for n, s in c["input"]]
        want = sorted((n, s) for n, s in c["expected"])
        for chunk in (1, g["chunk"]):
            got = msa.all_seq_affine_chunk(recs, S, g["gap_open"], g["gap_extend"], chunk)
            assert sorted((n, orc.bases_to_string(s)) for n, s in got) == want, chunk


def test_appendix_known_answers():  # SURVEY.md appendix A regression values
    S = orc.DEFAULT_SCORE_MATRIX
    rows = [("ACGT", "ACGT", 382, "4M"), ("ACGT", "CGT", -139, "1D3M"), ("ACGT", "ACG", -139, "3M1D"),
            ("CGT", "ACGT", -139, "1I3M"), ("ACG", "ACGT", -139, "3M1I"), ("AGT", "ACGT", -148, "1M1I2M"),
            ("ACT", "ACGT", -148, "2M1I1M"), ("CGCGCGCGCG", "CGCGCGTTTTCGCG", 480, "6M4I4M"),
            ("CGCGCGCGCG", "CGAAAACGCGTTTTCGCG", -40, "2M4I4M4I4M")]
    for a, b, sc, cg in rows:
        s, c = orc.affine_gap_highmem(bases(a), bases(b), S, -400, -30)
        assert (s, orc.print_cigar(c)) == (sc, cg)
    rows = [("ACGT", "ACGT", 382, "4M"), ("ACGT", "CGT", -139, "1D3M"), ("AA", "GGGAATT", -1968, "3I2M2I"),
            ("GGGAATT", "AA", -1968, "3D2M2D"), ("AGTACGT", "ACGTACG", -287, "1M1I5M1D"),
            ("CGCGCGCGCG", "CGAAAACGCGTTTTCGCG", -2440, "2M4I4M4I4M")]
    for a, b, sc, cg in rows:
        s, c = orc.const_gap_highmem(bases(a), bases(b), S, -430)
        assert (s, orc.print_cigar(c)) == (sc, cg)


def test_edge_cases_highmem():  # SURVEY.md 2b edge cases
    S = orc.DEFAULT_SCORE_MATRIX
    e = np.zeros(0, dtype=np.uint8)
    assert orc.affine_gap_highmem(e, e, S, -400, -30) == (0, [(0, 0)])
    assert orc.affine_gap_highmem(e, bases("ACG"), S, -400, -30) == (-400 - 90, [(3, 1)])
    assert orc.affine_gap_highmem(bases("ACG"), e, S, -400, -30) == (-400 - 90, [(3, 2)])
    assert orc.affine_gap_highmem(bases("ACG"), e, S, -400, -30, True) == (0, [(3, 2)])
    assert orc.const_gap_highmem(e, e, S, -430) == (0, [(0, 0)])
    assert orc.const_gap_highmem(e, bases("AC"), S, -430) == (-860, [(2, 1)])
    with pytest.raises(orc.OracleError) as ei:  # lowercase indexes past the 5x5 matrix -> Go panics
        orc.affine_gap_highmem(bases("acg"), bases("ACG"), S, -400, -30)
    assert ei.value.code == orc.ORC_EBASE
    with pytest.raises(orc.OracleError) as ei:  # AffineGapChunk log.Fatalf on ragged length
        orc.affine_gap_chunk(bases("ACGT"), bases("ACG"), S, -400, -30, 3)
    assert ei.value.code == orc.ORC_ECHUNK


def test_lowmem_equals_highmem_single_board_random():
    rng = np.random.default_rng(7)
    S = orc.HUMAN_CHIMP_TWO_SCORE_MATRIX
    for _ in range(300):
        n, m = int(rng.integers(1, 60)), int(rng.integers(1, 60))
        a = rng.integers(0, 4, n, dtype=np.uint8)
        b = rng.integers(0, 4, m, dtype=np.uint8)
        assert orc.affine_gap_lowmem(a, b, S, -600, -150) == orc.affine_gap_highmem(a, b, S, -600, -150)
        assert orc.const_gap_lowmem(a, b, S, -430) == orc.const_gap_highmem(a, b, S, -430)


def test_batch_threads_match_single():
    rng = np.random.default_rng(11)
    S = orc.HUMAN_CHIMP_TWO_SCORE_MATRIX
    al = [rng.integers(0, 4, int(rng.integers(0, 40)), dtype=np.uint8) for _ in range(64)]
    be = [rng.integers(0, 4, int(rng.integers(0, 40)), dtype=np.uint8) for _ in range(64)]
    ao = np.concatenate([[0], np.cumsum([len(x) for x in al])]).astype(np.int64)
    bo = np.concatenate([[0], np.cumsum([len(x) for x in be])]).astype(np.int64)
    ac, bc = np.concatenate(al), np.concatenate(be)
    for mode in (0, 1, 2):
        sc, off, cg = orc.batch(ac, ao, bc, bo, S, -600, -150, mode, True, 3)
        for p in range(64):
            if mode == 2:
                s, c = orc.const_gap_highmem(al[p], be[p], S, -600)
            else:
                s, c = orc.affine_gap_highmem(al[p], be[p], S, -600, -150, mode == 1)
            got = [(int(r), int(o)) for r, o in cg[off[p]:off[p + 1]]]
            assert (int(sc[p]), got) == (s, c), (mode, p)


# ---- gsw extend step: no reference test asserts these two functions ("parity unpinned"),
# so the C oracle is cross-checked against a second, line-by-line Python transcription of the Go loops.
def _py_extend(alpha, beta, S, g, left):
    n, m = len(alpha), len(beta)
    M = [[0] * (m + 1) for _ in range(n + 1)]
    T = [[""] * (m + 1) for _ in range(n + 1)]

    def tmt(a, b, c):  # cigar.TripleMaxTrace (cigar/tools.go:58-66)
        if a >= b and a >= c:
            return a, "M"
        if b >= c:
            return b, "I"
        return c, "D"

    cur_max, mi, mj = 0, 0, 0
    if left:  # genomeGraph/search.go:236-251
        for i in range(1, n + 1):
            for j in range(1, m + 1):
                M[i][j], T[i][j] = tmt(M[i - 1][j - 1] + int(S[alpha[i - 1]][beta[j - 1]]), M[i][j - 1] + g,
                                       M[i - 1][j] + g)
                if M[i][j] < 0:
                    M[i][j] = 0
    else:  # :280-299
        for i in range(n + 1):
            for j in range(m + 1):
                if i == 0 and j == 0:
                    M[i][j] = 0
                elif i == 0:
                    M[i][j], T[i][j] = M[i][j - 1] + g, "I"
                elif j == 0:
                    M[i][j], T[i][j] = M[i - 1][j] + g, "D"
                else:
                    M[i][j], T[i][j] = tmt(M[i - 1][j - 1] + int(S[alpha[i - 1]][beta[j - 1]]), M[i][j - 1] + g,
                                           M[i - 1][j] + g)
                if M[i][j] > cur_max:
                    cur_max, mi, mj = M[i][j], i, j
    route = []
    i, j = (n, m) if left else (mi, mj)
    while (M[i][j] > 0) if left else (i > 0 or j > 0):
        op = T[i][j]
        if route and route[-1][1] == op:
            route[-1] = (route[-1][0] + 1, op)
        else:
            route.append((1, op))
        if op == "M":
            i, j = i - 1, j - 1
        elif op == "I":
            j -= 1
        else:
            i -= 1
    return (M[n][m], route, i, j) if left else (M[mi][mj], route, mi, mj)


def test_extend_oracle_against_python_transcription():
    rng = np.random.default_rng(234)
    S = orc.HUMAN_CHIMP_TWO_SCORE_MATRIX
    cases = [(np.zeros(0, np.uint8), np.zeros(0, np.uint8)), (bases("ACGT"), np.zeros(0, np.uint8)),
             (np.zeros(0, np.uint8), bases("ACGT")), (bases("ACGTACGT"), bases("ACGTACGT")),
             (bases("AAAAAAAA"), bases("AAAA")), (bases("ACACACACAC"), bases("CACACA"))]
    for _ in range(150):
        n, m = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        a, b = random_pair(rng, n, m, identity=float(rng.choice([0.6, 0.9, 1.0])), alphabet=int(rng.choice([2, 4, 5])))
        cases.append((a, b))
    for a, b in cases:
        for g in (-600, -100, 0):
            assert orc.left_dynamic_aln(a, b, S, g) == _py_extend(a, b, S, g, True), (a, b, g)
            assert orc.right_dynamic_aln(a, b, S, g) == _py_extend(a, b, S, g, False), (a, b, g)


def test_ungapped_diagonal_route_theorem():
    """The screening pass of the checkpoint path (ckpt_classify_kernel) relies on: in freeEndGaps mode, if the score
    equals the score of the ungapped diagonal that ends where the traceback leaves the last column (r*, m), the route
    is D x (r* - m), M x m, D x (n - r*).  Checked here against the pinned oracle on random, mutated, tie-heavy and
    unrelated pairs (CPU only)."""
    rng = np.random.default_rng(99)
    hits = 0
    for trial in range(1500):
        n, m = int(rng.integers(20, 220)), int(rng.integers(1, 60))
        kind = trial % 4
        if kind == 0:
            a, b = random_pair(rng, n, m, identity=float(rng.choice([0.7, 0.9, 1.0])))
        elif kind == 1:  # exact window, maybe with substitutions only
            a = rng.integers(0, 4, size=n, dtype=np.uint8)
            s = int(rng.integers(0, n - m + 1)) if n >= m else 0
            b = a[s:s + m].copy() if n >= m else rng.integers(0, 4, size=m, dtype=np.uint8)
            for k in rng.integers(0, m, size=int(rng.integers(0, 3))):
                b[k] = (b[k] + 1) % 4
        elif kind == 2:  # repeats: ties everywhere
            unit = rng.integers(0, 4, size=int(rng.integers(1, 4)), dtype=np.uint8)
            a, b = np.resize(unit, n).astype(np.uint8), np.resize(unit, m).astype(np.uint8)
        else:
            a, b = rng.integers(0, 4, size=n, dtype=np.uint8), rng.integers(0, 4, size=m, dtype=np.uint8)
        S = orc.HUMAN_CHIMP_TWO_SCORE_MATRIX if trial % 2 else orc.DEFAULT_SCORE_MATRIX
        O, E = ((-600, -150), (-400, -30), (0, -30))[trial % 3]
        score, cig = orc.affine_gap_highmem(a, b, S, O, E, free_end_gaps=True)
        tail_d = cig[-1][0] if cig and cig[-1][1] == 2 and len(cig) > 1 else 0  # D x (n - r*) closes the route
        rs = n - tail_d
        if rs < m:
            continue
        diag = int(sum(int(S[a[rs - m + k], b[k]]) for k in range(m)))
        if diag != score:
            continue
        hits += 1
        want = ([(rs - m, 2)] if rs > m else []) + [(m, 0)] + ([(n - rs, 2)] if n > rs else [])
        assert cig == want, (trial, n, m, cig, want)
    assert hits > 300
