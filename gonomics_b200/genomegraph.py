"""Host-side mirror of the gsw extend step of gonomics' `genomeGraph` package over the libgnxalign C ABI.

    LeftDynamicAln     genomeGraph/search.go:234-274
    RightDynamicAln    genomeGraph/search.go:276-321
    IndexGenomeIntoMap genomeGraph/index.go:21-44      (SeedIndex; nodes without edges)
    seedMapMemPool     genomeGraph/search.go:567-602   (SeedIndex.seeds_for_reads + heapSortSeeds)
    heapSortSeeds      genomeGraph/search.go:339-373   (host-side ordering of the returned seeds)

Both are linear-gap DPs over `cigar.Cigar{RunLength int; Op byte}` (Op in 'M','I','D', cigar/cigar.go:15-18)
whose route is returned in TRACEBACK order (the callers reverse it, search.go:196,230).  The reference
passes a scratch matrix and a dynamicScoreKeeper; neither carries state into the call (resetDynamicScore
takes its argument by value, search.go:104-107, and the callers hand in an empty route), so they are
accepted and ignored here.  Every call goes through the CUDA library; there is no CPU path.
"""
from __future__ import annotations

from typing import List, NamedTuple, Optional, Sequence, Tuple

import numpy as np

import ctypes as C

from ._lib import GNX_ECAP, GNX_EXT_LEFT, GNX_EXT_RIGHT, SEED_DTYPE, GnxError
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


def LeftLocal(alpha, beta, scores, gapPen: int = -600, m=None, trace=None, ctx=None):
    """genomeGraph.LeftLocal (genomeGraph/localAlignment.go:95-140): (score, route, minI, maxI, minJ, maxJ); ops '=','X','I','D'."""
    from ._lib import GNX_EXT_LEFT_LOCAL
    sc, route, i, j = extend_pairs(GNX_EXT_LEFT_LOCAL, [alpha], [beta], scores, gapPen, ctx)[0]
    return sc, route, i, len(alpha), j, len(beta)


def RightLocal(alpha, beta, scores, gapPen: int = -600, m=None, trace=None, ctx=None):
    """genomeGraph.RightLocal (genomeGraph/localAlignment.go:142-196): (score, route, 0, maxI, 0, maxJ)."""
    from ._lib import GNX_EXT_RIGHT_LOCAL
    sc, route, i, j = extend_pairs(GNX_EXT_RIGHT_LOCAL, [alpha], [beta], scores, gapPen, ctx)[0]
    return sc, route, 0, i, 0, j


# ---- the perfect-match seed step ---------------------------------------------------
class SeedDev(NamedTuple):
    """genomeGraph.SeedDev (genomeGraph/index.go:11-19); NextPart is always nil for edge-less nodes."""
    TargetId: int
    TargetStart: int
    QueryStart: int
    Length: int
    PosStrand: bool
    TotalLength: int


class SeedIndex:
    """genomeGraph.IndexGenomeIntoMap(genome, seedLen, seedStep) for nodes without edges, resident on the GPU
    together with the nodes' TwoBit encoding."""

    def __init__(self, nodes: Sequence[np.ndarray], seedLen: int, seedStep: int, ctx: Optional[Context] = None):
        self.ctx = ctx or default_context()
        self._L = self.ctx._L
        cat, off = _concat(nodes)
        h = C.c_void_p(None)
        self._h = None
        self.ctx._check(self._L.gnx_seed_index_new(self.ctx._h, cat.ctypes.data, off.ctypes.data, len(off) - 1, int(seedLen),
                                                   int(seedStep), C.byref(h)))
        self._h = h
        n = C.c_int64(0)
        self._L.gnx_seed_index_info(self._h, C.byref(n))
        self.n_entries, self.seedLen, self.seedStep = n.value, seedLen, seedStep

    def close(self):
        if getattr(self, "_h", None) and getattr(self.ctx, "_h", None):
            self._L.gnx_seed_index_free(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def entries(self) -> Tuple[np.ndarray, np.ndarray]:
        """(keys, locations) sorted by key; a key's locations are in the reference's insertion order."""
        key = np.zeros(max(self.n_entries, 1), dtype=np.uint64)
        loc = np.zeros(max(self.n_entries, 1), dtype=np.uint64)
        self.ctx._check(self._L.gnx_seed_index_download(self.ctx._h, self._h, key.ctypes.data, loc.ctypes.data))
        return key[:self.n_entries], loc[:self.n_entries]

    def seed_batch(self, reads_cat: np.ndarray, read_off: np.ndarray, out=None) -> Tuple[np.ndarray, np.ndarray]:
        """seedMapMemPool's seeds of every read in APPEND order: (seeds[SEED_DTYPE], offsets [n_reads+1]).
        `out` = (seeds buffer, offsets buffer) lets the caller provide (e.g. page-locked) result arrays."""
        cat = np.ascontiguousarray(reads_cat, dtype=np.uint8)
        off = np.ascontiguousarray(read_off, dtype=np.int64)
        n = len(off) - 1
        if out is not None:
            seeds, soff = out
            rc = self._L.gnx_seed_batch(self.ctx._h, self._h, cat.ctypes.data, off.ctypes.data, n, seeds.ctypes.data,
                                        soff.ctypes.data, len(seeds))
            self.ctx._check(rc)
            return seeds[:int(soff[n])], soff
        soff = np.zeros(n + 1, dtype=np.int64)
        cap = max(8 * n, 64)
        while True:
            seeds = np.zeros(cap, dtype=SEED_DTYPE)
            rc = self._L.gnx_seed_batch(self.ctx._h, self._h, cat.ctypes.data, off.ctypes.data, n, seeds.ctypes.data,
                                        soff.ctypes.data, cap)
            if rc == GNX_ECAP:
                cap = int(soff[-1])
                continue
            self.ctx._check(rc)
            return seeds[:int(soff[-1])], soff


def gsw_batch(index: "SeedIndex", reads_cat: np.ndarray, read_off: np.ndarray, scores, paired: bool = False,
              cigar_cap: Optional[int] = None) -> Tuple[np.ndarray, np.ndarray]:
    """genomeGraph.GraphSmithWatermanToGiraf (WrapPairGiraf when paired) for every read of a block: gnx_gsw_batch.
    Returns (records[GIRAF_DTYPE], cigars[CIGAR_DTYPE]); record r's cigar is cigars[cigar_off : cigar_off + n_cigar]
    (n_cigar = -1: nil), ops are the bytes 'M','I','D','S'."""
    from ._lib import CIGAR_DTYPE, GIRAF_DTYPE
    cat = np.ascontiguousarray(reads_cat, dtype=np.uint8)
    off = np.ascontiguousarray(read_off, dtype=np.int64)
    S = np.ascontiguousarray(scores, dtype=np.int64)
    n = len(off) - 1
    recs = np.zeros(max(n, 1), dtype=GIRAF_DTYPE)
    cap = int(cigar_cap or 4 * n + 64)
    need = C.c_int64(0)
    while True:
        cig = np.zeros(max(cap, 1), dtype=CIGAR_DTYPE)
        rc = index._L.gnx_gsw_batch(index.ctx._h, index._h, cat.ctypes.data, off.ctypes.data, n, S.ctypes.data, int(S.shape[0]),
                                    int(bool(paired)), recs.ctypes.data, cig.ctypes.data, cap, C.byref(need))
        if rc == GNX_ECAP and need.value > cap:
            cap = int(need.value)
            continue
        index.ctx._check(rc)
        return recs[:n], cig[:int(need.value)]


def heapSortSeeds(a: List[SeedDev]) -> None:
    """genomeGraph.heapSortSeeds (search.go:339-373): in-place min-heap sort, i.e. descending TotalLength with
    the reference's (unstable) order among equal lengths."""
    def heapify(size, i):
        while True:
            l, r = 2 * i + 1, 2 * i + 2
            m = l if l < size and a[l].TotalLength < a[i].TotalLength else i
            if r < size and a[r].TotalLength < a[m].TotalLength:
                m = r
            if m == i:
                return
            a[i], a[m] = a[m], a[i]
            i = m
    for i in range(len(a) // 2 - 1, -1, -1):
        heapify(len(a), i)
    size = len(a)
    for i in range(len(a) - 1, 0, -1):
        a[0], a[i] = a[i], a[0]
        size -= 1
        heapify(size, 0)


def seedMapMemPool(index: SeedIndex, reads: Sequence[np.ndarray]) -> List[List[SeedDev]]:
    """genomeGraph.seedMapMemPool for a batch of reads: the GPU enumerates and extends the seeds, the final
    ordering (search.go:596-600) is applied here.  Lists longer than 100 use sort.Slice in the reference (an
    unstable pdqsort whose order among equal TotalLength is unspecified); a stable descending sort is used."""
    cat, off = _concat(reads)
    seeds, soff = index.seed_batch(cat, off)
    out = []
    for r in range(len(reads)):
        lst = [SeedDev(int(s["target_id"]), int(s["target_start"]), int(s["query_start"]), int(s["length"]),
                       bool(s["pos_strand"]), int(s["total_length"])) for s in seeds[soff[r]:soff[r + 1]]]
        if len(lst) > 100:
            lst.sort(key=lambda x: -x.TotalLength)
        else:
            heapSortSeeds(lst)
        out.append(lst)
    return out
