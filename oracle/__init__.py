"""ctypes bindings for the CPU ORACLE (test infrastructure, NOT the product).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  ``gonomics_b200`` never does.

The oracle restates the Go reference (``/root/reference/align``) in plain C
(``gnx_oracle.c``); see ``gnx_oracle.h`` for the file:line map and parity status.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libgnxoracle.so")

ORC_OK, ORC_EBASE, ORC_ECAP, ORC_ECHUNK, ORC_EPANIC, ORC_EUNDEF, ORC_ENOMEM = range(7)


class OrcCigar(C.Structure):
    _fields_ = [("run_length", C.c_int64), ("op", C.c_uint8)]


CIGAR_DTYPE = np.dtype({"names": ["run_length", "op"], "formats": ["<i8", "u1"], "offsets": [0, 8], "itemsize": 16})

# align/align.go:28-64 -- the four score matrices, [alpha][beta] over A,C,G,T,N
DEFAULT_SCORE_MATRIX = np.array(
    [[91, -114, -31, -123, -44], [-114, 100, -125, -31, -43], [-31, -125, 100, -114, -43],
     [-123, -31, -114, 91, -44], [-44, -43, -43, -44, -43]], dtype=np.int64)
HOXD55_SCORE_MATRIX = np.array(
    [[91, -114, -31, -123, 0], [-114, 100, -125, -31, 0], [-31, -125, 100, -114, 0],
     [-123, -31, -114, 91, 0], [0, 0, 0, 0, 0]], dtype=np.int64)
MOUSE_RAT_SCORE_MATRIX = HOXD55_SCORE_MATRIX.copy()
HUMAN_CHIMP_TWO_SCORE_MATRIX = np.array(
    [[90, -330, -236, -356, -208], [-330, 100, -318, -236, -196], [-236, -318, 100, -330, -196],
     [-356, -236, -330, 90, -208], [-208, -196, -196, -208, -202]], dtype=np.int64)

_BASE_OF = {"A": 0, "C": 1, "G": 2, "T": 3, "N": 4, "a": 5, "c": 6, "g": 7, "t": 8, "n": 9, "-": 10, ".": 11, "*": 12}
_RUNE_OF = "ACGTNacgtn-.*"


def string_to_bases(s: str) -> np.ndarray:
    """dna.StringToBases (dna/convert.go): A,C,G,T,N,a,c,g,t,n,-,.,* -> 0..12."""
    return np.fromiter((_BASE_OF[ch] for ch in s), dtype=np.uint8, count=len(s))


def bases_to_string(b: Sequence[int]) -> str:
    return "".join(_RUNE_OF[int(x)] for x in b)


def to_upper(b: np.ndarray) -> np.ndarray:
    """dna.AllToUpper: lowercase a,c,g,t,n (5..9) -> 0..4."""
    b = np.asarray(b, dtype=np.uint8).copy()
    low = (b >= 5) & (b <= 9)
    b[low] -= 5
    return b


def build() -> str:
    """Compile the oracle (gcc, a second or two).  Building the checker is not using it."""
    subprocess.run(["make", "-s", "-C", _HERE], check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < max(
                os.path.getmtime(os.path.join(_HERE, f)) for f in ("gnx_oracle.c", "gnx_twobit_oracle.c", "gnx_oracle.h")):
            build()
        L = C.CDLL(_SO)
        u8p, i64p, cgp = C.POINTER(C.c_uint8), C.POINTER(C.c_int64), C.POINTER(OrcCigar)
        i64, ci = C.c_int64, C.c_int
        L.orc_affine_highmem.argtypes = [u8p, i64, u8p, i64, i64p, ci, i64, i64, ci, ci, i64p, cgp, i64, i64p]
        L.orc_const_highmem.argtypes = [u8p, i64, u8p, i64, i64p, ci, i64, ci, i64p, cgp, i64, i64p]
        L.orc_affine_lowmem.argtypes = [u8p, i64, u8p, i64, i64p, ci, i64, i64, i64, i64, i64p, cgp, i64, i64p]
        L.orc_const_lowmem.argtypes = [u8p, i64, u8p, i64, i64p, ci, i64, i64, i64, i64p, cgp, i64, i64p]
        L.orc_affine_chunk.argtypes = [u8p, i64, u8p, i64, i64p, ci, i64, i64, i64, i64p, cgp, i64, i64p]
        L.orc_multi_affine_chunk.argtypes = [u8p, i64, i64, u8p, i64, i64, i64p, ci, i64, i64, i64, i64p, cgp, i64, i64p]
        L.orc_left_dynamic_aln.argtypes = [u8p, i64, u8p, i64, i64p, ci, i64, i64p, cgp, i64, i64p, i64p, i64p]
        L.orc_right_dynamic_aln.argtypes = [u8p, i64, u8p, i64, i64p, ci, i64, i64p, cgp, i64, i64p, i64p, i64p]
        L.orc_left_dynamic_aln.restype = ci
        L.orc_right_dynamic_aln.restype = ci
        L.orc_batch.argtypes = [u8p, i64p, u8p, i64p, i64, i64p, ci, i64, i64, ci, ci, ci, i64p, cgp, i64p, i64p]
        for f in ("orc_affine_highmem", "orc_const_highmem", "orc_affine_lowmem", "orc_const_lowmem",
                  "orc_affine_chunk", "orc_multi_affine_chunk", "orc_batch"):
            getattr(L, f).restype = ci
        u64p, u32p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)
        L.orc_new_twobit.argtypes = [u8p, i64, ci, u64p]
        L.orc_new_twobit.restype = ci
        L.orc_get_base.argtypes = [u64p, C.c_uint64]
        L.orc_get_base.restype = C.c_uint8
        for f in ("orc_count_right", "orc_count_left"):
            getattr(L, f).argtypes = [u64p, i64, u64p, i64, i64, i64]
            getattr(L, f).restype = i64
        L.orc_seed_index.argtypes = [u8p, i64p, i64, ci, ci, u64p, u64p, i64]
        L.orc_seed_index.restype = i64
        L.orc_seeds_for_read.argtypes = [u64p, u64p, i64, u64p, i64p, i64p, u8p, u8p, i64, ci, u32p, i64]
        L.orc_seeds_for_read.restype = i64
        _lib = L
    return _lib


class OracleError(RuntimeError):
    def __init__(self, code: int, what: str):
        super().__init__(f"{what}: oracle status {code}")
        self.code = code


def _u8(a) -> Tuple[np.ndarray, "C._Pointer"]:
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(C.POINTER(C.c_uint8))


def _i64(a) -> Tuple[np.ndarray, "C._Pointer"]:
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(C.POINTER(C.c_int64))


Cigar = List[Tuple[int, int]]  # [(RunLength, Op)]


def _run(fn, what, alpha, beta, scores, pre, post, cap) -> Tuple[int, Cigar]:
    a, ap = _u8(alpha)
    b, bp = _u8(beta)
    s, sp = _i64(scores)
    dim = int(s.shape[0])
    score = C.c_int64(0)
    n_out = C.c_int64(0)
    buf = np.zeros(max(cap, 1), dtype=CIGAR_DTYPE)
    rc = fn(ap, len(a), bp, len(b), sp, dim, *pre, *post, C.byref(score),
            buf.ctypes.data_as(C.POINTER(OrcCigar)), cap, C.byref(n_out))
    if rc != ORC_OK:
        raise OracleError(rc, what)
    k = n_out.value
    return score.value, [(int(buf["run_length"][i]), int(buf["op"][i])) for i in range(k)]


def affine_gap_highmem(alpha, beta, scores, gap_open, gap_extend, free_end_gaps=False, want_cigar=True):
    """align.AffineGap_highMem (free_end_gaps=False) / align.AffineGapLocal (True)."""
    cap = len(alpha) + len(beta) + 2
    return _run(lib().orc_affine_highmem, "affine_highmem", alpha, beta, scores,
                (int(gap_open), int(gap_extend), int(bool(free_end_gaps)), int(bool(want_cigar))), (), cap)


def affine_gap_local(target, query, scores, gap_open, gap_extend):
    return affine_gap_highmem(target, query, scores, gap_open, gap_extend, True)


def const_gap_highmem(alpha, beta, scores, gap_pen, want_cigar=True):
    cap = len(alpha) + len(beta) + 2
    return _run(lib().orc_const_highmem, "const_highmem", alpha, beta, scores,
                (int(gap_pen), int(bool(want_cigar))), (), cap)


def affine_gap_lowmem(alpha, beta, scores, gap_open, gap_extend, ci=10000, cj=10000):
    """align.AffineGap (ci=cj=10000) / align.AffineGap_customizeCheckersize."""
    cap = len(alpha) + len(beta) + 2
    return _run(lib().orc_affine_lowmem, "affine_lowmem", alpha, beta, scores,
                (int(gap_open), int(gap_extend), int(ci), int(cj)), (), cap)


def const_gap_lowmem(alpha, beta, scores, gap_pen, ci=10000, cj=10000):
    """align.ConstGap (ci=cj=10000) / align.ConstGap_customizeCheckersize."""
    cap = len(alpha) + len(beta) + 2
    return _run(lib().orc_const_lowmem, "const_lowmem", alpha, beta, scores, (int(gap_pen), int(ci), int(cj)), (), cap)


def affine_gap_chunk(alpha, beta, scores, gap_open, gap_extend, chunk):
    cap = len(alpha) + len(beta) + 2
    return _run(lib().orc_affine_chunk, "affine_chunk", alpha, beta, scores,
                (int(gap_open), int(gap_extend), int(chunk)), (), cap)


def multi_affine_gap_chunk(group_a: np.ndarray, group_b: np.ndarray, scores, gap_open, gap_extend, chunk=1):
    """multipleAffineGap (chunk=1) / multipleAffineGapChunk over 2-D uint8 groups [n_seq, len]."""
    ga = np.ascontiguousarray(group_a, dtype=np.uint8)
    gb = np.ascontiguousarray(group_b, dtype=np.uint8)
    s, sp = _i64(scores)
    cap = ga.shape[1] + gb.shape[1] + 2
    buf = np.zeros(cap, dtype=CIGAR_DTYPE)
    score, n_out = C.c_int64(0), C.c_int64(0)
    rc = lib().orc_multi_affine_chunk(ga.ctypes.data_as(C.POINTER(C.c_uint8)), ga.shape[0], ga.shape[1],
                                      gb.ctypes.data_as(C.POINTER(C.c_uint8)), gb.shape[0], gb.shape[1],
                                      sp, int(s.shape[0]), int(gap_open), int(gap_extend), int(chunk),
                                      C.byref(score), buf.ctypes.data_as(C.POINTER(OrcCigar)), cap, C.byref(n_out))
    if rc != ORC_OK:
        raise OracleError(rc, "multi_affine_chunk")
    return score.value, [(int(buf["run_length"][i]), int(buf["op"][i])) for i in range(n_out.value)]


def _extend(fn, what, alpha, beta, scores, gap_pen):
    a, ap = _u8(alpha)
    b, bp = _u8(beta)
    s, sp = _i64(scores)
    cap = len(a) + len(b) + 2
    buf = np.zeros(cap, dtype=CIGAR_DTYPE)
    score, n_out, ri, rj = C.c_int64(0), C.c_int64(0), C.c_int64(0), C.c_int64(0)
    rc = fn(ap, len(a), bp, len(b), sp, int(s.shape[0]), int(gap_pen), C.byref(score),
            buf.ctypes.data_as(C.POINTER(OrcCigar)), cap, C.byref(n_out), C.byref(ri), C.byref(rj))
    if rc != ORC_OK:
        raise OracleError(rc, what)
    route = [(int(buf["run_length"][k]), chr(int(buf["op"][k]))) for k in range(n_out.value)]
    return score.value, route, ri.value, rj.value


def left_dynamic_aln(alpha, beta, scores, gap_pen):
    """genomeGraph.LeftDynamicAln (genomeGraph/search.go:234): (score, route in traceback order, i, j)."""
    return _extend(lib().orc_left_dynamic_aln, "left_dynamic_aln", alpha, beta, scores, gap_pen)


def right_dynamic_aln(alpha, beta, scores, gap_pen):
    """genomeGraph.RightDynamicAln (genomeGraph/search.go:276): (score, route in traceback order, maxI, maxJ)."""
    return _extend(lib().orc_right_dynamic_aln, "right_dynamic_aln", alpha, beta, scores, gap_pen)


def batch(alpha_cat, alpha_off, beta_cat, beta_off, scores, gap_open, gap_extend, mode, want_cigar=True, n_threads=1):
    """Threaded batch of affineGap_highMem (mode 0 global / 1 free-end) or ConstGap_highMem (mode 2).

    Returns (scores[int64], cigar_off[int64 n+1], cigars[CIGAR_DTYPE]) with cigars compacted."""
    a, ap = _u8(alpha_cat)
    b, bp = _u8(beta_cat)
    ao, aop = _i64(alpha_off)
    bo, bop = _i64(beta_off)
    s, sp = _i64(scores)
    n_pairs = len(ao) - 1
    out_score = np.zeros(n_pairs, dtype=np.int64)
    if want_cigar:
        caps = (np.diff(ao) + np.diff(bo) + 1).astype(np.int64)
        off = np.zeros(n_pairs + 1, dtype=np.int64)
        np.cumsum(caps, out=off[1:])
        buf = np.zeros(max(int(off[-1]), 1), dtype=CIGAR_DTYPE)
        cnt = np.zeros(n_pairs, dtype=np.int64)
        offp = off.ctypes.data_as(C.POINTER(C.c_int64))
        bufp = buf.ctypes.data_as(C.POINTER(OrcCigar))
        cntp = cnt.ctypes.data_as(C.POINTER(C.c_int64))
    else:
        off = buf = cnt = None
        offp = bufp = cntp = None
    rc = lib().orc_batch(ap, aop, bp, bop, n_pairs, sp, int(s.shape[0]), int(gap_open), int(gap_extend), int(mode),
                         int(bool(want_cigar)), int(n_threads), out_score.ctypes.data_as(C.POINTER(C.c_int64)),
                         bufp, offp, cntp)
    if rc != ORC_OK:
        raise OracleError(rc, "batch")
    if not want_cigar:
        return out_score, None, None
    coff = np.zeros(n_pairs + 1, dtype=np.int64)
    np.cumsum(cnt, out=coff[1:])
    idx = np.repeat(off[:-1] - coff[:-1], cnt) + np.arange(int(coff[-1]), dtype=np.int64)
    return out_score, coff, buf[idx]


# ---- align/view.go restated for the golden tests (pretty printers only) ----
def print_cigar(cig: Cigar) -> str:
    """align.PrintCigar (align/view.go:25-33)."""
    return "".join(f"{r}{'MID'[op]}" for r, op in cig)


def view(alpha, beta, cig: Cigar) -> str:
    """align.View (align/view.go:37-63)."""
    one, two, i, j = [], [], 0, 0
    for run, op in cig:
        for _ in range(run):
            if op == 0:
                one.append(_RUNE_OF[int(alpha[i])]); two.append(_RUNE_OF[int(beta[j])]); i += 1; j += 1
            elif op == 1:
                one.append("-"); two.append(_RUNE_OF[int(beta[j])]); j += 1
            else:
                one.append(_RUNE_OF[int(alpha[i])]); two.append("-"); i += 1
    return "".join(one) + "\n" + "".join(two) + "\n"


# ---- dna/dnaTwoBit and the perfect-match seed step (gnx_twobit_oracle.c) ----
def _u64(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a, a.ctypes.data_as(C.POINTER(C.c_uint64))


def new_twobit(seq, lead: int = 0) -> Tuple[np.ndarray, int]:
    """dnaTwoBit.NewTwoBit (lead=0) / NewTwoBitRainbow(seq)[lead]: (Seq []uint64, Len)."""
    s, sp = _u8(seq)
    total = len(s) + lead
    out = np.zeros(max((total + 31) // 32, 1), dtype=np.uint64)
    rc = lib().orc_new_twobit(sp, len(s), int(lead), out.ctypes.data_as(C.POINTER(C.c_uint64)))
    if rc != ORC_OK:
        raise OracleError(rc, "new_twobit")
    return out[:(total + 31) // 32], total


def get_base(words, pos: int) -> int:
    """dnaTwoBit.GetBase."""
    w, wp = _u64(words)
    return int(lib().orc_get_base(wp, int(pos)))


def count_right_matches(one, one_len, start_one, two, two_len, start_two) -> int:
    """dnaTwoBit.CountRightMatches; -1 = Fatalf (different offsets), -2 = index-out-of-range panic."""
    a, ap = _u64(one)
    b, bp = _u64(two)
    return int(lib().orc_count_right(ap, int(one_len), bp, int(two_len), int(start_one), int(start_two)))


def count_left_matches(one, one_len, start_one, two, two_len, start_two) -> int:
    """dnaTwoBit.CountLeftMatches; same error returns."""
    a, ap = _u64(one)
    b, bp = _u64(two)
    return int(lib().orc_count_left(ap, int(one_len), bp, int(two_len), int(start_one), int(start_two)))


def reverse_complement(seq) -> np.ndarray:
    """dna.ReverseComplement (dna/modify.go:72,111-115)."""
    comp = np.array([3, 2, 1, 0, 4, 8, 7, 6, 5, 9, 10, 11, 12], dtype=np.uint8)  # dna/modify.go:72 complementArray
    return comp[np.asarray(seq, dtype=np.uint8)[::-1]].copy()


def seed_index(genome_cat, node_off, seed_len: int, seed_step: int):
    """genomeGraph.IndexGenomeIntoMap for edge-less nodes as a (key, loc)-sorted array."""
    g, gp = _u8(genome_cat)
    no, nop = _i64(node_off)
    cap = max(int(len(g) // max(seed_step, 1)) + len(no) + 1, 1)
    key = np.zeros(cap, dtype=np.uint64)
    loc = np.zeros(cap, dtype=np.uint64)
    k = lib().orc_seed_index(gp, nop, len(no) - 1, int(seed_len), int(seed_step),
                             key.ctypes.data_as(C.POINTER(C.c_uint64)), loc.ctypes.data_as(C.POINTER(C.c_uint64)), cap)
    if k < 0:
        raise OracleError(int(k), "seed_index")
    return key[:k].copy(), loc[:k].copy()


def pack_nodes(genome_cat, node_off):
    """NewTwoBit of every node: (words concatenated, word offsets) -- what the reference keeps in Node.SeqTwoBit."""
    g = np.ascontiguousarray(genome_cat, dtype=np.uint8)
    no = np.ascontiguousarray(node_off, dtype=np.int64)
    n_nodes = len(no) - 1
    wo = np.zeros(n_nodes + 1, dtype=np.int64)
    np.cumsum((np.diff(no) + 31) // 32, out=wo[1:])
    words = np.zeros(max(int(wo[-1]), 1), dtype=np.uint64)
    for k in range(n_nodes):
        w, _ = new_twobit(g[no[k]:no[k + 1]])
        words[wo[k]:wo[k + 1]] = w
    return words, wo


def seeds_for_read(idx_key, idx_loc, genome_cat, node_off, read, seed_len: int, packed=None) -> np.ndarray:
    """genomeGraph.seedMapMemPool for one read, seeds in append order: uint32 [n, 6] =
    (TargetId, TargetStart, QueryStart, Length, PosStrand, TotalLength).  `packed` = pack_nodes(...) (cached)."""
    no, nop = _i64(node_off)
    words, wo = packed if packed is not None else pack_nodes(genome_cat, no)
    ik, ikp = _u64(idx_key)
    il, ilp = _u64(idx_loc)
    r, rp = _u8(read)
    rc_, rcp = _u8(reverse_complement(r))
    cap = 64
    while True:
        out = np.zeros((cap, 6), dtype=np.uint32)
        k = lib().orc_seeds_for_read(ikp, ilp, len(ik), words.ctypes.data_as(C.POINTER(C.c_uint64)),
                                     wo.ctypes.data_as(C.POINTER(C.c_int64)), nop, rp, rcp, len(r), int(seed_len),
                                     out.ctypes.data_as(C.POINTER(C.c_uint32)), cap)
        if k == -1:
            cap *= 4
            continue
        if k < 0:
            raise OracleError(int(k), "seeds_for_read")
        return out[:k].copy()
