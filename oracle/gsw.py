"""TEST INFRASTRUCTURE (oracle): a sequential restatement of gonomics' gsw per-read driver for a LINEAR genome graph
(nodes without edges), read after read and seed after seed exactly as the reference loops.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may import this; the product (gnx_gsw_batch) never does.

    GraphSmithWatermanToGiraf   genomeGraph/toGiraf.go:17-72
    WrapPairGiraf, setGirafFlags, isProperPairAlign, getGirafFlags   :117-137, :171-203
    seedCouldBeBetter           genomeGraph/index.go:102-121
    LeftAlignTraversal / RightAlignTraversal base cases   genomeGraph/search.go:169-181, :206-216
    getLeftTargetBases / getRightBases   :133-145
    perfectMatchBig, scoreSeedSeq        genomeGraph/align.go:73-87
    cigar.Append / Concat / AppendSoftClips / QueryLength   cigar/tools.go:4-40, cigar/cigar.go:145-156
    heapSortSeeds               genomeGraph/search.go:339-373 (lists <= 100; longer ones: sort.Slice, order among equal
                                TotalLength unspecified -- a stable descending sort here and in the product)

Facts of the Go code this restatement depends on (read from the source; the Go toolchain is absent, so they are
PARITY-UNPINNED by a run of the reference -- the reference has no asserting test for this function either):
  * scoreKeeper and dynamicScoreKeeper travel BY VALUE (toGiraf.go:17; search.go:104-107,124-133): the reset
    functions are no-ops, every GraphSmithWatermanToGiraf call starts from the worker's zero-valued keeper
    (routines.go:18-56: `dynamicScoreKeeper{}`, route == nil), so Left/RightDynamicAln always build their route in a
    fresh slice -- no aliasing between the left and right routes;
  * on a node without edges LeftAlignTraversal / RightAlignTraversal return the DP routes unreversed, i.e. in
    TRACEBACK order (the ReverseCigar calls sit on the branching path only, search.go:196,230), and the paths they
    return are empty (`AddPath`'s result is dropped, :175), so Path.Nodes is the seed's node;
  * a seed spanning the whole read skips the extension and keeps the PREVIOUS iteration's left/right alignments and
    queryEnd in the keeper (toGiraf.go:45-49,58,62): restated literally (observable only through stale state, which
    the descending seed order keeps empty in practice).
"""
from __future__ import annotations

from typing import List, NamedTuple, Optional, Sequence, Tuple

import numpy as np

import oracle as orc


class Giraf(NamedTuple):
    QStart: int
    QEnd: int
    PosStrand: bool
    TStart: int
    TEnd: int
    Nodes: Tuple[int, ...]
    Cigar: Optional[Tuple[Tuple[int, str], ...]]  # None == Go nil
    AlnScore: int
    Flag: int


def seed_could_be_better(seedLen, currBestScore, perfectScore, queryLen, maxMatch, minMatch, leastSevereMismatch,
                         leastSevereMatchMismatchChange) -> bool:
    """genomeGraph/index.go:102-121 (Go integer division truncates; all operands are non-negative here)."""
    seeds = queryLen // (seedLen + 1)
    remainder = queryLen % (seedLen + 1)
    if seedLen * maxMatch >= currBestScore and perfectScore - ((queryLen - seedLen) * minMatch) >= currBestScore:
        return True
    if (seedLen * seeds * maxMatch + seeds * leastSevereMismatch >= currBestScore and
            perfectScore - remainder * minMatch + seeds * leastSevereMatchMismatchChange >= currBestScore):
        return True
    if (seedLen * seeds * maxMatch + remainder * maxMatch + (seeds + 1) * leastSevereMismatch >= currBestScore and
            perfectScore + (seeds + 1) * leastSevereMatchMismatchChange >= currBestScore):
        return True
    return False


def heap_sort_seeds(a: List[Sequence[int]]) -> None:
    """genomeGraph.heapSortSeeds (search.go:339-373) on rows whose column 5 is TotalLength."""
    def heapify(size, i):
        while True:
            l, r = 2 * i + 1, 2 * i + 2
            m = l if l < size and a[l][5] < a[i][5] else i
            if r < size and a[r][5] < a[m][5]:
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


def cigar_append(alpha: list, beta: Tuple[int, str]) -> list:
    """cigar.Append (cigar/tools.go:4-11)."""
    if alpha and alpha[-1][1] == beta[1]:
        alpha[-1] = (alpha[-1][0] + beta[0], beta[1])
    else:
        alpha.append(beta)
    return alpha


def cigar_concat(alpha: list, beta: list) -> list:
    """cigar.Concat (:14-23)."""
    if not alpha:
        return beta
    if beta:
        alpha = cigar_append(alpha, beta[0])
        beta = beta[1:]
    return alpha + beta


def append_soft_clips(front: int, length_of_read: int, cigars: list) -> list:
    """cigar.AppendSoftClips (:26-40), including its quirk: with front > 0 and nothing left to clip at the end the body
    is dropped and only the leading soft clip is returned."""
    curr = sum(r for r, op in cigars if op in "MIS=X")  # cigar.QueryLength / ConsumesQuery
    if front == 0 and curr >= length_of_read:
        return cigars
    answer = []
    if front > 0:
        answer.append((front, "S"))
    if front + curr < length_of_read:
        answer = answer + cigars + [(length_of_read - front - curr, "S")]
    return answer


class LinearGenome:
    """The pieces of a GenomeGraph without edges the driver touches: node sequences, their TwoBit form and the seed map."""

    def __init__(self, nodes: Sequence[np.ndarray], seed_len: int, seed_step: int):
        self.nodes = [np.ascontiguousarray(n, dtype=np.uint8) for n in nodes]
        self.cat = np.concatenate(self.nodes + [np.zeros(0, dtype=np.uint8)])
        self.off = np.zeros(len(self.nodes) + 1, dtype=np.int64)
        np.cumsum([len(n) for n in self.nodes], out=self.off[1:])
        self.seed_len = seed_len
        self.key, self.loc = orc.seed_index(self.cat, self.off, seed_len, seed_step)
        self.packed = orc.pack_nodes(self.cat, self.off)


def graph_smith_waterman_to_giraf(gg: LinearGenome, read: np.ndarray, scores: np.ndarray) -> Giraf:
    """genomeGraph.GraphSmithWatermanToGiraf (toGiraf.go:17-72) for one read; Flag = getGirafFlags (:187-196)."""
    read = np.ascontiguousarray(read, dtype=np.uint8)
    read_rc = orc.reverse_complement(read)
    S = np.asarray(scores, dtype=np.int64)
    best = dict(QStart=0, QEnd=0, PosStrand=True, TStart=0, TEnd=0, Nodes=(), Cigar=None, AlnScore=0)
    perfect = int(sum(int(S[b][b]) for b in read))
    extension = perfect // 600 + len(read)
    hits = [tuple(int(x) for x in row) for row in orc.seeds_for_read(gg.key, gg.loc, gg.cat, gg.off, read, gg.seed_len, gg.packed)]
    if len(hits) > 100:
        hits.sort(key=lambda s: -s[5])  # SortSeedLen: sort.Slice (unspecified among equals; stable here)
    else:
        heap_sort_seeds(hits)
    # scoreKeeper state that survives from one seed to the next (by-value keeper inside ONE call)
    left_aln: list = []
    right_aln: list = []
    query_end = 0
    for (tid, tstart, qstart, length, pos, total) in hits:
        if not seed_could_be_better(total, best["AlnScore"], perfect, len(read), 100, 90, -196, -296):
            break
        curr = read if pos else read_rc
        # tailSeed = the seed itself (NextPart == nil without edges)
        seed_score = int(sum(int(S[b][b]) for b in curr[qstart:qstart + length]))
        if total == len(curr):
            target_start, target_end, query_start = tstart, tstart + length, qstart
            curr_score = seed_score
        else:
            node = gg.nodes[tid]
            ext = extension - total
            # LeftAlignTraversal base case: s.Seq = n.Seq[refEnd - min(refEnd, ext) : refEnd], seq == nil
            ref_end = tstart
            lt = node[ref_end - min(ref_end, ext):ref_end]
            lscore, lroute, li, lj = orc.left_dynamic_aln(lt, curr[:qstart], S, -600)
            left_aln = [(int(r), str(o)) for r, o in lroute]
            target_start = ref_end - len(lt) + li
            query_start = lj
            # RightAlignTraversal base case: s.Seq = n.Seq[start : start + min(len(n.Seq) - start, ext)]
            start = tstart + length
            rt = node[start:start + min(len(node) - start, ext)]
            rscore, rroute, ri, rj = orc.right_dynamic_aln(rt, curr[qstart + length:], S, -600)
            right_aln = [(int(r), str(o)) for r, o in rroute]
            target_end = ri + start
            query_end = rj
            curr_score = lscore + seed_score + rscore
        if curr_score > best["AlnScore"]:
            cig = cigar_concat(cigar_append(list(left_aln), (total, "M")), list(right_aln))
            best = dict(QStart=query_start, QEnd=qstart + query_start + query_end + total - 1, PosStrand=bool(pos),
                        TStart=target_start, TEnd=target_end, Nodes=(tid,),
                        Cigar=tuple(append_soft_clips(query_start, len(curr), cig)), AlnScore=curr_score)
    flag = (4 if best["PosStrand"] else 0) + (2 if best["AlnScore"] < 1200 else 0)
    return Giraf(Flag=flag, **best)


def wrap_pair_giraf(gg: LinearGenome, fwd: np.ndarray, rev: np.ndarray, scores) -> Tuple[Giraf, Giraf]:
    """genomeGraph.WrapPairGiraf + setGirafFlags (toGiraf.go:117-137): Fwd.Flag += 8 + 16 + 16 as written there."""
    f, r = graph_smith_waterman_to_giraf(gg, fwd, scores), graph_smith_waterman_to_giraf(gg, rev, scores)
    ff, rf = f.Flag + 8 + 16 + 16, r.Flag
    proper = False
    if abs(f.TStart - r.TStart) < 10000:
        if f.TStart < r.TStart and f.PosStrand and not r.PosStrand:
            proper = True
        if f.TStart > r.TStart and not f.PosStrand and r.PosStrand:
            proper = True
    if proper:
        ff += 1
        rf += 1
    return f._replace(Flag=ff & 0xff), r._replace(Flag=rf & 0xff)  # Flag is a uint8
