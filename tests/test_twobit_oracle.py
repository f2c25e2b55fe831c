"""Pin the 2-bit / seed ORACLE (oracle/gnx_twobit_oracle.c) to the reference's known answers
(dna/dnaTwoBit/perfectAlign_test.go, dnaTwoBit_test.go -> tests/golden/twobit.json) and to independent
base-level restatements.  CPU only."""
import numpy as np

import oracle as orc
from golden_util import load


def naive_pack(seq, lead=0):
    """word = OR_i base_i << (62 - 2i) mod 2^64 over the `lead` x A + seq clone (dnaTwoBit.go:28-42)."""
    clone = [0] * lead + [int(x) for x in seq]
    words = []
    for w in range((len(clone) + 31) // 32):
        v = 0
        for i, b in enumerate(clone[32 * w:32 * w + 32]):
            v |= (b << (62 - 2 * i)) & 0xFFFFFFFFFFFFFFFF
        words.append(v)
    return np.array(words, dtype=np.uint64), len(clone)


def test_count_matches_known_answers():  # perfectAlign_test.go:97-110 TestCounting
    g = load("twobit")
    assert len(g["count_cases"]) == 4
    for c in g["count_cases"]:
        a, la = orc.new_twobit(orc.string_to_bases(c["seq_a"]))
        b, lb = orc.new_twobit(orc.string_to_bases(c["seq_b"]))
        assert orc.count_left_matches(a, la, c["start_a"], b, lb, c["start_b"]) == c["left"], c["names"]
        assert orc.count_right_matches(a, la, c["start_a"], b, lb, c["start_b"]) == c["right"], c["names"]


def test_get_base_known_answers():  # dnaTwoBit_test.go:17-42 TestDnaToFromString
    g = load("twobit")
    for s in g["get_base_strings"]:
        w, _ = orc.new_twobit(orc.string_to_bases(s))
        for pos, base in g["get_base_checks"]:
            assert orc.get_base(w, pos) == "ACGT".index(base), (s, pos)


def test_pack_matches_base_level_restatement():
    rng = np.random.default_rng(5)
    for n in [0, 1, 31, 32, 33, 64, 150, 257]:
        for hi in (4, 13):  # clean ACGT; N / lowercase / gap codes that spill into neighbouring bases
            seq = rng.integers(0, hi, size=n, dtype=np.uint8)
            for lead in (0, 1, 17, 31):
                w, ln = orc.new_twobit(seq, lead)
                nw, nl = naive_pack(seq, lead)
                assert ln == nl and np.array_equal(w, nw), (n, hi, lead)
                if hi == 4:
                    assert all(orc.get_base(w, lead + i) == seq[i] for i in range(0, n, 7))


def test_count_matches_against_base_loops():  # perfectAlign_test.go:81-95 currentMethodRight / currentMethodLeft
    rng = np.random.default_rng(6)
    for _ in range(300):
        n, m = int(rng.integers(1, 200)), int(rng.integers(1, 200))
        a = rng.integers(0, 4, size=n, dtype=np.uint8)
        b = a[:m].copy() if m <= n else np.concatenate([a, rng.integers(0, 4, size=m - n, dtype=np.uint8)])
        for k in rng.integers(0, min(n, m), size=3):
            b[k] = (b[k] + 1) % 4
        sa = int(rng.integers(0, min(n, m)))
        sb = sa % 32 + 32 * int(rng.integers(0, (min(n, m) - 1 - sa % 32) // 32 + 1)) if min(n, m) > sa % 32 else sa
        wa, la = orc.new_twobit(a)
        wb, lb = orc.new_twobit(b)
        right = 0
        while sa + right < n and sb + right < m and a[sa + right] == b[sb + right]:
            right += 1
        assert orc.count_right_matches(wa, la, sa, wb, lb, sb) == right
        left = 0
        while sa - left >= 0 and sb - left >= 0 and a[sa - left] == b[sb - left]:
            left += 1
        # the word loop keeps going while whole words match, exactly like the base loop, and stops at word 0
        assert orc.count_left_matches(wa, la, sa, wb, lb, sb) == left
    w, ln = orc.new_twobit(np.zeros(40, dtype=np.uint8))
    assert orc.count_right_matches(w, ln, 3, w, ln, 4) == -1  # different offsets: log.Fatalf
    assert orc.count_left_matches(w, ln, 64, w, ln, 0) == -2  # Seq[2] of a 2-word sequence: panic


def naive_seeds(nodes, read, seed_len, seed_step):
    """Base-level restatement of seedMapMemPool for clean (A,C,G,T) sequences and edge-less nodes."""
    index = {}
    for ni, node in enumerate(nodes):
        for pos in range(0, len(node) - seed_len + 1, seed_step):
            index.setdefault(bytes(node[pos:pos + seed_len]), []).append((ni, pos))
    out = []
    rc = orc.reverse_complement(read)
    for start in range(len(read) - seed_len + 1):
        for strand, q in ((1, read), (0, rc)):
            for ni, pos in index.get(bytes(q[start:start + seed_len]), []):
                node = nodes[ni]
                left = 0
                while start - left >= 0 and pos - left >= 0 and q[start - left] == node[pos - left]:
                    left += 1
                rs, ns = start - (left - 1), pos - (left - 1)
                right = 0
                while rs + right < len(q) and ns + right < len(node) and q[rs + right] == node[ns + right]:
                    right += 1
                out.append((ni, ns, rs, right, strand, right))
    return np.array(out, dtype=np.uint32).reshape(-1, 6)


def test_seed_oracle_against_base_level_restatement():
    rng = np.random.default_rng(7)
    nodes = [rng.integers(0, 4, size=n, dtype=np.uint8) for n in (700, 64, 31, 1500)]
    nodes[3][100:400] = nodes[0][50:350]  # a repeat: k-mers with several locations
    cat = np.concatenate(nodes)
    off = np.cumsum([0] + [len(x) for x in nodes]).astype(np.int64)
    for seed_len, seed_step in ((32, 32), (20, 8), (11, 1)):
        key, loc = orc.seed_index(cat, off, seed_len, seed_step)
        assert np.all(key[:-1] <= key[1:])
        for trial in range(25):
            ni = int(rng.integers(0, 4))
            if len(nodes[ni]) < 60:
                continue
            s = int(rng.integers(0, len(nodes[ni]) - 50))
            read = nodes[ni][s:s + int(rng.integers(40, 160))].copy()
            for k in rng.integers(0, len(read), size=int(rng.integers(0, 4))):
                read[k] = (read[k] + 1) % 4
            if trial % 2:
                read = orc.reverse_complement(read)
            got = orc.seeds_for_read(key, loc, cat, off, read, seed_len)
            want = naive_seeds(nodes, read, seed_len, seed_step)
            assert np.array_equal(got, want), (seed_len, seed_step, trial)
