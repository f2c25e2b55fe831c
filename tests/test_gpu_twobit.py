"""Parity of the device 2-bit encoding, CountRight/LeftMatches and the perfect-match seed step
against the oracle and the reference's known answers, through the C ABI.  Bit-exact."""
import numpy as np
import pytest

import oracle as orc
from golden_util import load
from gonomics_b200 import _lib, align, dnatwobit, genomegraph

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = align.Context(0)
    yield c
    c.close()


def test_golden_count_matches_and_get_base(ctx):  # perfectAlign_test.go TestCounting, dnaTwoBit_test.go TestDnaToFromString
    g = load("twobit")
    for c in g["count_cases"]:
        a = dnatwobit.NewTwoBit(orc.string_to_bases(c["seq_a"]), ctx)
        b = dnatwobit.NewTwoBit(orc.string_to_bases(c["seq_b"]), ctx)
        assert dnatwobit.CountLeftMatches(a, c["start_a"], b, c["start_b"]) == c["left"], c["names"]
        assert dnatwobit.CountRightMatches(a, c["start_a"], b, c["start_b"]) == c["right"], c["names"]
    for s in g["get_base_strings"]:
        frag = dnatwobit.NewTwoBit(orc.string_to_bases(s), ctx)
        assert frag.Len == len(s)
        for pos, base in g["get_base_checks"]:
            assert dnatwobit.GetBase(frag, pos) == "ACGT".index(base), (s, pos)


def oracle_set(seqs, lead):
    words, off, lens = [], [0], []
    for s in seqs:
        w, ln = orc.new_twobit(s, lead)
        words.append(w)
        off.append(off[-1] + len(w))
        lens.append(ln)
    return np.concatenate(words + [np.zeros(0, dtype=np.uint64)]), np.array(off, dtype=np.int64), np.array(lens, dtype=np.int64)


@pytest.mark.parametrize("alphabet", [4, 13])
def test_pack_ragged_batches_all_leads(ctx, alphabet):
    """NewTwoBit / NewTwoBitRainbow[lead] of ragged batches (empty sequences, every alignment of the byte window,
    bases > 3 whose raw byte spills into the neighbouring bases' bits)."""
    rng = np.random.default_rng(11 + alphabet)
    lens = [0, 1, 2, 31, 32, 33, 63, 64, 65, 150, 150, 151, 0, 1000, 7, 96] + [int(x) for x in rng.integers(0, 400, size=60)]
    seqs = [rng.integers(0, alphabet, size=n, dtype=np.uint8) for n in lens]
    for lead in (0, 1, 2, 15, 16, 30, 31):
        tb = dnatwobit.TwoBitSet.from_seqs(seqs, lead, ctx)
        words, woff, ln = tb.download()
        ow, ooff, oln = oracle_set(seqs, lead)
        assert np.array_equal(woff, ooff) and np.array_equal(ln, oln)
        bad = np.nonzero(words != ow)[0]
        assert len(bad) == 0, (lead, bad[:5], [hex(int(words[i])) for i in bad[:3]], [hex(int(ow[i])) for i in bad[:3]])
        if alphabet == 4:  # GetBase round trip (only defined for 2-bit-clean input)
            back, boff = tb.unpack()
            for k, s in enumerate(seqs):
                assert np.array_equal(back[boff[k] + lead:boff[k + 1]], s)
                assert not back[boff[k]:boff[k] + lead].any()  # the lead is dna.A
        tb.close()


def test_pack_uniform_reads_and_long_sequence(ctx):
    rng = np.random.default_rng(12)
    reads = rng.integers(0, 4, size=(5000, 150), dtype=np.uint8)
    tb = dnatwobit.TwoBitSet(reads.reshape(-1), np.arange(5001, dtype=np.int64) * 150, 0, ctx)
    words, woff, ln = tb.download()
    ow, ooff, oln = oracle_set(list(reads[:300]) + list(reads[-5:]), 0)
    assert np.array_equal(words[:300 * 5], ow[:300 * 5]) and np.array_equal(words[-25:], ow[-25:])
    back, _ = tb.unpack()
    assert np.array_equal(back, reads.reshape(-1))
    tb.close()
    # one long sequence (the 16-byte fast path) with an N island and a ragged tail; full-size round trip
    n = 3_000_017
    genome = rng.integers(0, 4, size=n, dtype=np.uint8)
    tb = dnatwobit.TwoBitSet.from_seqs([genome], 0, ctx)
    back, _ = tb.unpack()
    assert np.array_equal(back, genome)
    genome[1000:1100] = 4
    genome[2_000_000] = 9
    tb2 = dnatwobit.TwoBitSet.from_seqs([genome], 0, ctx)
    words, _, _ = tb2.download()
    ow, _ = orc.new_twobit(genome)
    assert np.array_equal(words, ow)
    q = rng.integers(0, n, size=1000)
    assert np.array_equal(tb.get_bases(np.zeros(1000, dtype=np.int64), q),
                          np.array([orc.get_base(tb.download()[0], int(p)) for p in q], dtype=np.uint8))
    tb.close()
    tb2.close()


def test_count_matches_random_queries_and_errors(ctx):
    rng = np.random.default_rng(13)
    ones, twos = [], []
    for _ in range(40):
        n = int(rng.integers(1, 500))
        a = rng.integers(0, 4, size=n, dtype=np.uint8)
        b = a.copy()
        for k in rng.integers(0, n, size=int(rng.integers(0, 5))):
            b[k] = (b[k] + 1) % 4
        cut = int(rng.integers(0, 40))
        ones.append(a)
        twos.append(b[cut - cut % 32:] if rng.random() < 0.3 and n > cut else b)
    one = dnatwobit.TwoBitSet.from_seqs(ones, 0, ctx)
    two = dnatwobit.TwoBitSet.from_seqs(twos, 0, ctx)
    ow = [orc.new_twobit(s) for s in ones]
    tw = [orc.new_twobit(s) for s in twos]
    q1, s1, q2, s2 = [], [], [], []
    for _ in range(4000):
        i, j = int(rng.integers(0, 40)), int(rng.integers(0, 40))
        a = int(rng.integers(0, len(ones[i])))
        cand = [b for b in range(a % 32, len(twos[j]), 32)]
        if not cand:
            continue
        q1.append(i), s1.append(a), q2.append(j), s2.append(cand[int(rng.integers(0, len(cand)))])
    for d, fn in ((_lib.GNX_MATCH_RIGHT, orc.count_right_matches), (_lib.GNX_MATCH_LEFT, orc.count_left_matches)):
        got = dnatwobit.count_matches(d, one, two, q1, s1, q2, s2)
        want = np.array([fn(ow[i][0], ow[i][1], a, tw[j][0], tw[j][1], b) for i, a, j, b in zip(q1, s1, q2, s2)])
        assert np.array_equal(got, want), np.nonzero(got != want)[0][:5]
    # errors: the first offending query in order decides, as for a sequential caller
    with pytest.raises(_lib.GnxError) as e:
        dnatwobit.count_matches(_lib.GNX_MATCH_RIGHT, one, two, [0, 0, 0], [0, 1, 0], [0, 0, 0], [0, 2, 10 ** 6])
    assert e.value.code == _lib.GNX_EOFFSET and "element 1" in str(e.value)
    with pytest.raises(_lib.GnxError) as e:
        dnatwobit.count_matches(_lib.GNX_MATCH_LEFT, one, two, [0, 0], [0, 0], [0, 0], [32 * 1000, 1])
    assert e.value.code == _lib.GNX_EINDEX and "element 0" in str(e.value)
    with pytest.raises(_lib.GnxError) as e:
        one.get_bases([0], [32 * 1000])
    assert e.value.code == _lib.GNX_EINDEX


def make_genome(rng):
    nodes = [rng.integers(0, 4, size=n, dtype=np.uint8) for n in (5000, 64, 31, 12000, 33)]
    nodes[3][1000:1900] = nodes[0][200:1100]      # a repeat: k-mers with several locations, across nodes
    nodes[3][5000:5064] = 0                       # poly-A
    nodes[0][3000:3040] = 4                       # N island: windows touching it are not indexed
    nodes[3][7000] = 6                            # a lowercase base: its raw byte spills in key and TwoBit word
    return nodes


@pytest.mark.parametrize("seed_len,seed_step", [(32, 32), (20, 8), (11, 3)])
def test_seed_index_and_seeds_match_oracle(ctx, seed_len, seed_step):
    rng = np.random.default_rng(14)
    nodes = make_genome(rng)
    cat = np.concatenate(nodes)
    off = np.cumsum([0] + [len(x) for x in nodes]).astype(np.int64)
    ix = genomegraph.SeedIndex(nodes, seed_len, seed_step, ctx)
    okey, oloc = orc.seed_index(cat, off, seed_len, seed_step)
    key, loc = ix.entries()
    assert np.array_equal(key, okey) and np.array_equal(loc, oloc)
    reads = []
    for trial in range(120):
        ni = (0, 3)[trial % 2]
        s = int(rng.integers(0, len(nodes[ni]) - 200))
        read = nodes[ni][s:s + int(rng.integers(30, 200))].copy()
        for k in rng.integers(0, len(read), size=int(rng.integers(0, 4))):
            read[k] = (read[k] + 1) % 4 if read[k] < 4 else read[k]
        if trial % 3 == 0:
            read = orc.reverse_complement(read)
        if trial % 10 == 7:
            read[int(rng.integers(0, len(read)))] = 4  # an N in the read: the exact (non-funnel) rainbow path
        reads.append(read)
    reads += [np.zeros(0, dtype=np.uint8), np.zeros(5, dtype=np.uint8), np.zeros(150, dtype=np.uint8),
              rng.integers(0, 4, size=150, dtype=np.uint8), nodes[1].copy(), nodes[0][:seed_len].copy()]
    rcat, roff = align._concat(reads)
    seeds, soff = ix.seed_batch(rcat, roff)
    n_total = 0
    for r, read in enumerate(reads):
        want = orc.seeds_for_read(okey, oloc, cat, off, read, seed_len)
        got = seeds[soff[r]:soff[r + 1]]
        got = np.stack([got[f] for f in got.dtype.names], axis=1) if len(got) else np.zeros((0, 6), dtype=np.uint32)
        assert np.array_equal(got, want), (r, len(read), got[:4], want[:4])
        n_total += len(want)
    assert n_total > 200
    # host-side ordering: heapSortSeeds gives descending TotalLength
    out = genomegraph.seedMapMemPool(ix, reads[:20])
    for lst in out:
        assert all(lst[i].TotalLength >= lst[i + 1].TotalLength for i in range(len(lst) - 1))
    ix.close()


def test_seed_errors(ctx):
    with pytest.raises(_lib.GnxError) as e:
        genomegraph.SeedIndex([np.zeros(100, dtype=np.uint8)], 33, 32, ctx)  # index.go:22-24 log.Fatalf
    assert e.value.code == _lib.GNX_EARG
    ix = genomegraph.SeedIndex([np.zeros(100, dtype=np.uint8)], 16, 16, ctx)
    with pytest.raises(_lib.GnxError) as e:
        ix.seed_batch(np.array([0, 1, 13, 2] * 10, dtype=np.uint8), np.array([0, 40], dtype=np.int64))
    assert e.value.code == _lib.GNX_EBASE  # complementArray[13]: index out of range
    ix.close()
