"""Helpers shared by the oracle-golden tests and the GPU parity tests."""
import json
import os

import numpy as np

import oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

MATRICES = {
    "Default": orc.DEFAULT_SCORE_MATRIX,
    "HoxD55": orc.HOXD55_SCORE_MATRIX,
    "MouseRat": orc.MOUSE_RAT_SCORE_MATRIX,
    "HumanChimpTwo": orc.HUMAN_CHIMP_TWO_SCORE_MATRIX,
}


def load(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        return json.load(f)


def bases(s, upper=False):
    b = orc.string_to_bases(s)
    return orc.to_upper(b) if upper else b


def cigar_to_beds(aln, first_ins, first_del, chrom):
    """The BED derivation of cmd/cigarToBed/cigarToBed.go:93-129 (ins then del), as text."""
    ins, cur = [], first_ins - 1
    for i in range(len(aln) - 1):
        if aln[i][1] == 0 and aln[i + 1][1] == 1:
            start = cur + aln[i][0] + 1
            ins.append(f"{chrom}\t{start}\t{start + aln[i + 1][0]}\tins\n")
        if aln[i][1] != 2:
            cur += aln[i][0]
    dele, cur = [], first_del - 1
    for i in range(len(aln) - 1):
        if aln[i][1] == 0 and aln[i + 1][1] == 1:
            start = cur + aln[i][0]
            dele.append(f"{chrom}\t{start}\t{start + 1}\tdel\n")
        if aln[i][1] != 1:
            cur += aln[i][0]
    return "".join(ins), "".join(dele)


def random_pair(rng, n, m, identity=0.9, alphabet=4):
    """A related (alpha, beta) pair of exact lengths n, m with substitutions and indel bursts."""
    a = rng.integers(0, alphabet, size=n, dtype=np.uint8)
    src = a if n >= m else np.concatenate([a, rng.integers(0, alphabet, size=m - n, dtype=np.uint8)])
    start = int(rng.integers(0, max(len(src) - m, 0) + 1))
    out, i = [], start
    while len(out) < m and i < len(src):
        u = rng.random()
        if u < (1 - identity) * 0.2:  # insertion burst
            out.extend(rng.integers(0, alphabet, size=int(rng.integers(1, 6))).tolist())
        elif u < (1 - identity) * 0.4:  # deletion burst
            i += int(rng.integers(1, 6))
        elif u < (1 - identity):
            out.append(int(rng.integers(0, alphabet)))
            i += 1
        else:
            out.append(int(src[i]))
            i += 1
    while len(out) < m:
        out.append(int(rng.integers(0, alphabet)))
    return a, np.array(out[:m], dtype=np.uint8)
