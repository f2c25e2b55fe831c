"""Host-side mirror of the gsw extend step of gonomics' `genomeGraph` package over the libgnxalign C ABI.

    LeftDynamicAln     genomeGraph/search.go:234-274
    RightDynamicAln    genomeGraph/search.go:276-321

Both are linear-gap DPs over `cigar.Cigar{RunLength int; Op byte}` (Op in 'M','I','D', cigar/cigar.go:15-18)
whose route is returned in TRACEBACK order (the callers reverse it, search.go:196,230).  The reference
passes a scratch matrix and a dynamicScoreKeeper; neither carries state into the call (resetDynamicScore
takes its argument by value, search.go:104-107, and the callers hand in an empty route), so they are
accepted and ignored here.  Every call goes through the CUDA library; there is no CPU path.
"""
from __future__ import annotations

from typing import List, NamedTuple, Optional, Sequence, Tuple

import numpy as np

from ._lib import GNX_EXT_LEFT, GNX_EXT_RIGHT
from .align import Context, _concat, default_context


class Cigar(NamedTuple):
    """cigar.Cigar (cigar/cigar.go:21-24)."""
    RunLength: int
    Op: str


def extend_pairs(side: int, alphas: Sequence[np.ndarray], betas: Sequence[np.ndarray], scores, gapPen: int,
                 ctx: Optional[Context] = None) -> List[Tuple[int, List[Cigar], int, int]]:
    """Batched form: one GPU launch for all (alpha, beta) pairs; returns [(score, route, i, j)...]."""
    ctx = ctx or default_context()
    ac, ao = _concat(alphas)
    bc, bo = _concat(betas)
    sc, ei, ej, off, cig = ctx.extend_batch(side, ac, ao, bc, bo, scores, gapPen)
    rl, op = cig["run_length"], cig["op"]
    return [(int(sc[p]), [Cigar(int(rl[k]), chr(int(op[k]))) for k in range(int(off[p]), int(off[p + 1]))],
             int(ei[p]), int(ej[p])) for p in range(len(sc))]


def LeftDynamicAln(alpha, beta, scores, matrix=None, gapPen: int = -600, dynamicScore=None, ctx=None):
    """genomeGraph.LeftDynamicAln (genomeGraph/search.go:234): (score, route, i, j)."""
    return extend_pairs(GNX_EXT_LEFT, [alpha], [beta], scores, gapPen, ctx)[0]


def RightDynamicAln(alpha, beta, scores, matrix=None, gapPen: int = -600, dynamicScore=None, ctx=None):
    """genomeGraph.RightDynamicAln (genomeGraph/search.go:276): (score, route, maxI, maxJ)."""
    return extend_pairs(GNX_EXT_RIGHT, [alpha], [beta], scores, gapPen, ctx)[0]
