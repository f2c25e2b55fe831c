"""Parity of the per-read gsw driver (gnx_gsw_batch) against the sequential restatement of
genomeGraph.GraphSmithWatermanToGiraf / WrapPairGiraf in oracle/gsw.py, through the C ABI.  Field by field, bit-exact."""
import numpy as np
import pytest

import oracle as orc
from oracle import gsw as ogsw
from gonomics_b200 import align, genomegraph

pytestmark = pytest.mark.gpu


def _genome(rng):
    """A linear reference with a repeat family (reads with many equal-length seeds), an N island and short nodes."""
    unit = rng.integers(0, 4, size=700, dtype=np.uint8)
    n0 = rng.integers(0, 4, size=60_000, dtype=np.uint8)
    for p in (5_000, 21_000, 40_000, 52_000):  # four diverged copies of the repeat
        u = unit.copy()
        for k in rng.integers(0, 700, size=6):
            u[k] = (u[k] + 1) % 4
        n0[p:p + 700] = u
    n0[30_000:30_040] = 4
    n1 = rng.integers(0, 4, size=9_000, dtype=np.uint8)
    n2 = rng.integers(0, 4, size=400, dtype=np.uint8)
    return [n0, n1, n2]


def _reads(rng, nodes, n_reads, read_len=150):
    reads = []
    for t in range(n_reads):
        ni = int(rng.choice([0, 0, 0, 1, 2]))
        node = nodes[ni]
        L = min(read_len, len(node))
        s = int(rng.integers(0, len(node) - L + 1))
        if t % 7 == 0:  # from the repeat family
            s = int(rng.choice([5_000, 21_000, 40_000, 52_000])) + int(rng.integers(0, 500))
            node, L = nodes[0], read_len
        if t % 11 == 0:  # hugging a node end: clipped extension windows
            s = 0 if t % 22 == 0 else len(node) - L
        read = node[s:s + L].copy()
        kind = t % 5
        if kind == 1:      # substitutions
            for k in rng.integers(0, L, size=int(rng.integers(1, 5))):
                read[k] = (read[k] + 1 + int(rng.integers(0, 3))) % 4 if read[k] < 4 else read[k]
        elif kind == 2:    # a deletion in the read
            k, d = int(rng.integers(40, L - 40)), int(rng.integers(1, 4))
            read = np.concatenate([read[:k], read[k + d:], node[s + L:s + L + d]])[:L]
        elif kind == 3:    # an insertion in the read
            k, d = int(rng.integers(40, L - 40)), int(rng.integers(1, 4))
            read = np.concatenate([read[:k], rng.integers(0, 4, size=d, dtype=np.uint8), read[k:]])[:L]
        elif kind == 4 and t % 3 == 0:  # unrelated sequence: no seed or junk seeds
            read = rng.integers(0, 4, size=L, dtype=np.uint8)
        if t % 2 == 1:
            read = orc.reverse_complement(read)
        if t % 13 == 5:
            read[int(rng.integers(0, len(read)))] = 4  # N in the read
        reads.append(np.ascontiguousarray(read, dtype=np.uint8))
    return reads


def _check(recs, cig, want, r):
    g = recs[r]
    got_cig = None if g["n_cigar"] < 0 else tuple(
        (int(c["run_length"]), chr(int(c["op"]))) for c in cig[int(g["cigar_off"]):int(g["cigar_off"]) + int(g["n_cigar"])])
    got = (int(g["q_start"]), int(g["q_end"]), bool(g["pos_strand"]), int(g["t_start"]), int(g["t_end"]),
           () if g["node"] < 0 else (int(g["node"]),), got_cig, int(g["aln_score"]), int(g["flag"]))
    assert got == tuple(want), (r, got, tuple(want))


@pytest.mark.parametrize("seed_len,seed_step", [(32, 32), (20, 8)])
def test_gsw_single_reads_match_oracle(seed_len, seed_step):
    rng = np.random.default_rng(2026 + seed_len)
    nodes = _genome(rng)
    S = orc.HUMAN_CHIMP_TWO_SCORE_MATRIX
    gg = ogsw.LinearGenome(nodes, seed_len, seed_step)
    reads = _reads(rng, nodes, 330) + [np.zeros(0, dtype=np.uint8), nodes[2].copy(), nodes[1][:seed_len].copy(),
                                       np.full(150, 4, dtype=np.uint8)]
    with align.Context(0) as ctx:
        ix = genomegraph.SeedIndex(nodes, seed_len, seed_step, ctx)
        rcat, roff = align._concat(reads)
        recs, cig = genomegraph.gsw_batch(ix, rcat, roff, S)
        n_mapped = 0
        for r, read in enumerate(reads):
            want = ogsw.graph_smith_waterman_to_giraf(gg, read, S)
            _check(recs, cig, want, r)
            n_mapped += want.AlnScore >= 1200
        assert n_mapped > 200
        # a cigar buffer that is too small: GNX_ECAP reports the size, the wrapper retries
        recs2, cig2 = genomegraph.gsw_batch(ix, rcat, roff, S, cigar_cap=3)
        assert np.array_equal(recs2, recs) and np.array_equal(cig2, cig)
        ix.close()


def test_gsw_paired_reads_match_oracle():
    rng = np.random.default_rng(777)
    nodes = _genome(rng)
    S = orc.HUMAN_CHIMP_TWO_SCORE_MATRIX
    gg = ogsw.LinearGenome(nodes, 32, 32)
    pairs = []
    for t in range(120):  # fragments of 300-500 bases: mates on opposite strands (proper pairs) and a few odd ones
        node = nodes[0]
        s = int(rng.integers(0, len(node) - 600))
        frag = int(rng.integers(300, 500))
        fwd = node[s:s + 150].copy()
        rev = orc.reverse_complement(node[s + frag - 150:s + frag])
        if t % 4 == 1:
            fwd, rev = rev, fwd
        if t % 9 == 2:
            rev = orc.reverse_complement(rev)  # same strand: not a proper pair
        if t % 10 == 3:
            rev = rng.integers(0, 4, size=150, dtype=np.uint8)
        fwd[int(rng.integers(0, 150))] ^= 1
        pairs.append((np.ascontiguousarray(fwd), np.ascontiguousarray(rev)))
    reads = [x for p in pairs for x in p]
    with align.Context(0) as ctx:
        ix = genomegraph.SeedIndex(nodes, 32, 32, ctx)
        rcat, roff = align._concat(reads)
        recs, cig = genomegraph.gsw_batch(ix, rcat, roff, S, paired=True)
        proper = 0
        for p, (fwd, rev) in enumerate(pairs):
            wf, wr = ogsw.wrap_pair_giraf(gg, fwd, rev, S)
            _check(recs, cig, wf, 2 * p)
            _check(recs, cig, wr, 2 * p + 1)
            proper += wr.Flag & 1
        assert proper > 50
        ix.close()
