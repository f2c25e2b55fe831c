"""ORACLE (test infrastructure): progressive multiple alignment drivers restated from
align/multiAlign.go:11-78,112-153 on top of the C oracle's profile DP."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from . import multi_affine_gap_chunk

GAP = 10  # dna.Gap

Group = List[Tuple[str, np.ndarray]]  # [(name, bases)]


def merge_multiple_alignments(alpha: Group, beta: Group, route) -> Group:
    """mergeMultipleAlignments (align/multiAlign.go:112-153)."""
    total = sum(r for r, _ in route)
    out = [(name, np.empty(total, dtype=np.uint8)) for name, _ in alpha + beta]
    ac = bc = col = 0
    for run, op in route:
        for _ in range(run):
            for k, (_, seq) in enumerate(out):
                if k < len(alpha):
                    seq[col] = alpha[k][1][ac] if op in (0, 2) else GAP
                else:
                    seq[col] = beta[k - len(alpha)][1][bc] if op in (0, 1) else GAP
            if op == 0:
                ac, bc = ac + 1, bc + 1
            elif op == 1:
                bc += 1
            else:
                ac += 1
            col += 1
    return out


def _stack(g: Group) -> np.ndarray:
    return np.stack([s for _, s in g]).astype(np.uint8)


def all_seq_affine_chunk(records: Group, scores, gap_open, gap_extend, chunk=1) -> Group:
    """AllSeqAffine (chunk=1, multiAlign.go:59-66) / AllSeqAffineChunk (:70-78): greedy merge of the
    best-scoring group pair (strict >, first found wins), all N(N-1)/2 profile DPs per round."""
    groups: List[Group] = [[r] for r in records]
    while len(groups) > 1:
        best = None
        for x in range(len(groups) - 1):
            for y in range(x + 1, len(groups)):
                score, route = multi_affine_gap_chunk(_stack(groups[x]), _stack(groups[y]), scores,
                                                      gap_open, gap_extend, chunk)
                if best is None or score > best[0]:
                    best = (score, x, y, route)
        _, x, y, route = best
        groups[x] = merge_multiple_alignments(groups[x], groups[y], route)  # mergeFastaGroups :20-25
        groups[y] = groups[-1]
        groups.pop()
    return groups[0]
