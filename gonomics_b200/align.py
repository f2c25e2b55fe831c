"""Host-side mirror of gonomics' `align` package API over the libgnxalign C ABI.

Names, argument order and meaning follow the Go functions (reference paths relative to the
gonomics tree):

    AffineGap, AffineGap_customizeCheckersize      align/affineGap.go:59,73
    AffineGap_highMem, AffineGapLocal              align/affineGap_highMem.go:99,105
    GoAffineGapLocalEngine, TargetQueryPair        align/affineGap_highMem.go:110-125
    ConstGap, ConstGap_customizeCheckersize        align/constGap.go:13,73
    ConstGap_highMem                               align/constGap_highMem.go:11
    AffineGapChunk                                 align/affineGap_highMem.go:227
    multipleAffineGap, multipleAffineGapChunk      align/affineGap_highMem.go:272,308
    nearestGroups[Chunk], AllSeqAffine[Chunk],
    mergeMultipleAlignments                        align/multiAlign.go:27-78,112-153

Every call goes through the CUDA library; there is no CPU path here.  Single-pair functions are
thin wrappers over the batched entry points (`affine_gap_batch`, `const_gap_batch`), which are the
performant boundary: one call per 10^4..10^7 pairs.

Sequences are `dna.Base` byte arrays (A,C,G,T,N = 0..4); cigars are returned as lists of
`Cigar(RunLength, Op)` with Op in {ColM=0, ColI=1, ColD=2} (align/align.go:12-24).
"""
from __future__ import annotations

import ctypes as C
import queue
import threading
from dataclasses import dataclass, field
from typing import List, NamedTuple, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import CIGAR_DTYPE, GNX_ECAP, GNX_FREE_END, GNX_GLOBAL, GNX_OK, GnxError

ColM, ColI, ColD = 0, 1, 2


class Cigar(NamedTuple):
    RunLength: int
    Op: int


# align/align.go:28-64
DefaultScoreMatrix = np.array(
    [[91, -114, -31, -123, -44], [-114, 100, -125, -31, -43], [-31, -125, 100, -114, -43],
     [-123, -31, -114, 91, -44], [-44, -43, -43, -44, -43]], dtype=np.int64)
HoxD55ScoreMatrix = np.array(
    [[91, -114, -31, -123, 0], [-114, 100, -125, -31, 0], [-31, -125, 100, -114, 0],
     [-123, -31, -114, 91, 0], [0, 0, 0, 0, 0]], dtype=np.int64)
MouseRatScoreMatrix = HoxD55ScoreMatrix.copy()
HumanChimpTwoScoreMatrix = np.array(
    [[90, -330, -236, -356, -208], [-330, 100, -318, -236, -196], [-236, -318, 100, -330, -196],
     [-356, -236, -330, 90, -208], [-208, -196, -196, -208, -202]], dtype=np.int64)


def _addr(a: Optional[np.ndarray]) -> Optional[int]:
    return None if a is None else a.ctypes.data


class Context:
    """One gnx_ctx: a CUDA device, its streams and scratch.  Not thread-safe (one per thread)."""

    def __init__(self, device: int = 0, workspace_bytes: int = 0):
        self._L = _lib.load()
        self._h = self._L.gnx_create(int(device), int(workspace_bytes))
        if not self._h:
            raise GnxError(_lib.GNX_ECUDA, self._L.gnx_last_error(None).decode())
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._L.gnx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- helpers ------------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != GNX_OK:
            raise GnxError(rc, self._L.gnx_last_error(self._h).decode())

    def set_option(self, name: str, value: int):
        self._check(self._L.gnx_set_option(self._h, name.encode(), int(value)))

    @property
    def launch_count(self) -> int:
        return int(self._L.gnx_launch_count(self._h))

    def last_kernel_path(self) -> Tuple[int, int]:
        """(impl, flags) of the last batch call: see gnx_last_kernel_path in include/gnxalign.h."""
        impl, flags = C.c_int(0), C.c_int(0)
        self._check(self._L.gnx_last_kernel_path(self._h, C.byref(impl), C.byref(flags)))
        return impl.value, flags.value

    def last_fill_stats(self) -> Tuple[float, int, int]:
        ms, n, cells = C.c_double(0), C.c_int64(0), C.c_int64(0)
        self._check(self._L.gnx_last_fill_stats(self._h, C.byref(ms), C.byref(n), C.byref(cells)))
        return ms.value, n.value, cells.value

    # ---- batched entry points (host buffers) ------------------------------------------
    def _batch(self, kind, alpha_cat, alpha_off, beta_cat, beta_off, scores, gap_open, gap_extend, want_cigar,
               cigar_cap=None, out=None):
        alpha_cat = np.ascontiguousarray(alpha_cat, dtype=np.uint8)
        beta_cat = np.ascontiguousarray(beta_cat, dtype=np.uint8)
        alpha_off = np.ascontiguousarray(alpha_off, dtype=np.int64)
        beta_off = np.ascontiguousarray(beta_off, dtype=np.int64)
        scores = np.ascontiguousarray(scores, dtype=np.int64)
        dim = int(scores.shape[0])
        n_pairs = len(alpha_off) - 1
        assert len(beta_off) == n_pairs + 1 and scores.shape == (dim, dim)
        if out is not None:
            out_score, out_off, out_cig = out
        else:
            out_score = np.zeros(n_pairs, dtype=np.int64)
            out_off = np.zeros(n_pairs + 1, dtype=np.int64) if want_cigar else None
            if want_cigar:
                if cigar_cap is None:
                    cigar_cap = 16 * n_pairs + 64
                out_cig = np.zeros(max(int(cigar_cap), 1), dtype=CIGAR_DTYPE)
            else:
                out_cig = None
        cap = 0 if out_cig is None else len(out_cig)
        if kind == 2:
            rc = self._L.gnx_const_batch(self._h, _addr(alpha_cat), _addr(alpha_off), _addr(beta_cat), _addr(beta_off),
                                         n_pairs, _addr(scores), dim, int(gap_open), int(bool(want_cigar)),
                                         _addr(out_score), _addr(out_cig), _addr(out_off), cap)
        else:
            rc = self._L.gnx_affine_batch(self._h, _addr(alpha_cat), _addr(alpha_off), _addr(beta_cat),
                                          _addr(beta_off), n_pairs, _addr(scores), dim, int(gap_open),
                                          int(gap_extend), GNX_FREE_END if kind == 1 else GNX_GLOBAL,
                                          int(bool(want_cigar)), _addr(out_score), _addr(out_cig), _addr(out_off), cap)
        if rc == GNX_ECAP and out is None:
            total = int(out_off[-1])
            out_cig = np.zeros(max(total, 1), dtype=CIGAR_DTYPE)
            self._check(self._L.gnx_copy_last_cigars(self._h, _addr(out_cig), len(out_cig)))
            rc = GNX_OK
        self._check(rc)
        if want_cigar:
            return out_score, out_off, out_cig[:int(out_off[-1])] if out is None else out_cig
        return out_score, None, None

    def affine_gap_batch(self, alpha_cat, alpha_off, beta_cat, beta_off, scores, gap_open, gap_extend,
                         free_end_gaps=False, want_cigar=True, cigar_cap=None, out=None):
        """Batched AffineGap_highMem (free_end_gaps=False) / AffineGapLocal (True).

        Returns (scores int64[n], cigar_off int64[n+1] | None, cigars CIGAR_DTYPE[] | None)."""
        return self._batch(1 if free_end_gaps else 0, alpha_cat, alpha_off, beta_cat, beta_off, scores, gap_open,
                           gap_extend, want_cigar, cigar_cap, out)

    def affine_gap_batch_twobit(self, alpha_words, alpha_len, beta_words, beta_len, scores, gap_open, gap_extend,
                                free_end_gaps=False, want_cigar=True, cigar_cap=None, out=None, n_pairs=None):
        """gnx_affine_batch_twobit: the batch in dnaTwoBit form (uint64 words, tightly packed).  alpha_len / beta_len are
        int64 arrays (one TwoBit.Len per pair) or plain ints for a uniform batch (then n_pairs must be given)."""
        alpha_words = np.ascontiguousarray(alpha_words, dtype=np.uint64)
        beta_words = np.ascontiguousarray(beta_words, dtype=np.uint64)
        scores = np.ascontiguousarray(scores, dtype=np.int64)
        uniform = np.isscalar(alpha_len)
        if uniform:
            assert np.isscalar(beta_len) and n_pairs is not None
            la = lb = None
            ua, ub = int(alpha_len), int(beta_len)
        else:
            la = np.ascontiguousarray(alpha_len, dtype=np.int64)
            lb = np.ascontiguousarray(beta_len, dtype=np.int64)
            n_pairs, ua, ub = len(la), 0, 0
        if out is not None:
            out_score, out_off, out_cig = out
        else:
            out_score = np.zeros(n_pairs, dtype=np.int64)
            out_off = np.zeros(n_pairs + 1, dtype=np.int64) if want_cigar else None
            out_cig = np.zeros(max(int(cigar_cap or 16 * n_pairs + 64), 1), dtype=CIGAR_DTYPE) if want_cigar else None
        cap = 0 if out_cig is None else len(out_cig)
        rc = self._L.gnx_affine_batch_twobit(self._h, _addr(alpha_words), _addr(la), ua, _addr(beta_words), _addr(lb), ub,
                                             n_pairs, _addr(scores), int(scores.shape[0]), int(gap_open), int(gap_extend),
                                             GNX_FREE_END if free_end_gaps else GNX_GLOBAL, int(bool(want_cigar)),
                                             _addr(out_score), _addr(out_cig), _addr(out_off), cap)
        if rc == GNX_ECAP and out is None:
            out_cig = np.zeros(max(int(out_off[-1]), 1), dtype=CIGAR_DTYPE)
            rc = self._L.gnx_copy_last_cigars(self._h, _addr(out_cig), len(out_cig))
        self._check(rc)
        if want_cigar:
            return out_score, out_off, out_cig[:int(out_off[-1])] if out is None else out_cig
        return out_score, None, None

    def affine_gap_chunk_batch(self, alpha_cat, alpha_off, beta_cat, beta_off, scores, gap_open, gap_extend, chunk,
                               cigar_cap=None):
        """Batched AffineGapChunk (align/affineGap_highMem.go:227): DP over chunk-sized blocks."""
        alpha_cat = np.ascontiguousarray(alpha_cat, dtype=np.uint8)
        beta_cat = np.ascontiguousarray(beta_cat, dtype=np.uint8)
        alpha_off = np.ascontiguousarray(alpha_off, dtype=np.int64)
        beta_off = np.ascontiguousarray(beta_off, dtype=np.int64)
        scores = np.ascontiguousarray(scores, dtype=np.int64)
        n_pairs = len(alpha_off) - 1
        out_score = np.zeros(n_pairs, dtype=np.int64)
        out_off = np.zeros(n_pairs + 1, dtype=np.int64)
        out_cig = np.zeros(max(int(cigar_cap or 16 * n_pairs + 64), 1), dtype=CIGAR_DTYPE)
        rc = self._L.gnx_affine_chunk_batch(self._h, _addr(alpha_cat), _addr(alpha_off), _addr(beta_cat),
                                            _addr(beta_off), n_pairs, _addr(scores), int(scores.shape[0]),
                                            int(gap_open), int(gap_extend), int(chunk), _addr(out_score), _addr(out_cig),
                                            _addr(out_off), len(out_cig))
        if rc == GNX_ECAP:
            out_cig = np.zeros(max(int(out_off[-1]), 1), dtype=CIGAR_DTYPE)
            rc = self._L.gnx_copy_last_cigars(self._h, _addr(out_cig), len(out_cig))
        self._check(rc)
        return out_score, out_off, out_cig[:int(out_off[-1])]

    def const_gap_batch(self, alpha_cat, alpha_off, beta_cat, beta_off, scores, gap_pen, want_cigar=True,
                        cigar_cap=None, out=None):
        """Batched ConstGap_highMem."""
        return self._batch(2, alpha_cat, alpha_off, beta_cat, beta_off, scores, gap_pen, 0, want_cigar, cigar_cap, out)

    def extend_batch(self, side, alpha_cat, alpha_off, beta_cat, beta_off, scores, gap_pen, want_cigar=True,
                     cigar_cap=None, out=None):
        """Batched genomeGraph.LeftDynamicAln (side=1) / RightDynamicAln (side=2) (genomeGraph/search.go:234-321).
        Returns (scores, end_i, end_j, cigar_off | None, cigars | None); cigar ops are the bytes 'M','I','D' and
        the route is in traceback order, as in the reference."""
        alpha_cat = np.ascontiguousarray(alpha_cat, dtype=np.uint8)
        beta_cat = np.ascontiguousarray(beta_cat, dtype=np.uint8)
        alpha_off = np.ascontiguousarray(alpha_off, dtype=np.int64)
        beta_off = np.ascontiguousarray(beta_off, dtype=np.int64)
        scores = np.ascontiguousarray(scores, dtype=np.int64)
        n_pairs = len(alpha_off) - 1
        if out is not None:  # caller-provided (e.g. page-locked) result arrays
            out_score, end_i, end_j, out_off, out_cig = out
        else:
            out_score = np.zeros(n_pairs, dtype=np.int64)
            end_i = np.zeros(n_pairs, dtype=np.int64)
            end_j = np.zeros(n_pairs, dtype=np.int64)
            out_off = np.zeros(n_pairs + 1, dtype=np.int64) if want_cigar else None
            out_cig = np.zeros(max(int(cigar_cap or 16 * n_pairs + 64), 1), dtype=CIGAR_DTYPE) if want_cigar else None
        rc = self._L.gnx_extend_batch(self._h, int(side), _addr(alpha_cat), _addr(alpha_off), _addr(beta_cat),
                                      _addr(beta_off), n_pairs, _addr(scores), int(scores.shape[0]), int(gap_pen),
                                      int(bool(want_cigar)), _addr(out_score), _addr(end_i), _addr(end_j),
                                      _addr(out_cig), _addr(out_off), 0 if out_cig is None else len(out_cig))
        if rc == GNX_ECAP and out is None:
            out_cig = np.zeros(max(int(out_off[-1]), 1), dtype=CIGAR_DTYPE)
            rc = self._L.gnx_copy_last_cigars(self._h, _addr(out_cig), len(out_cig))
        self._check(rc)
        if want_cigar:
            return out_score, end_i, end_j, out_off, out_cig[:int(out_off[-1])]
        return out_score, end_i, end_j, None, None

    def multi_affine_chunk_batch(self, groups, pair_x, pair_y, scores, gap_open, gap_extend, chunk=1,
                                 want_cigar=True, cigar_cap=None):
        """Batched multipleAffineGap (chunk=1) / multipleAffineGapChunk (align/affineGap_highMem.go:272-353):
        DP p aligns the sub-alignment groups[pair_x[p]] against groups[pair_y[p]].  groups: 2-D uint8 arrays
        (sequences x columns; dna.Base incl. lowercase 5..9 and Gap = 10)."""
        mats = [np.ascontiguousarray(g, dtype=np.uint8).reshape(len(g), -1) for g in groups]
        nseq = np.array([m.shape[0] for m in mats], dtype=np.int64)
        goff = np.zeros(len(mats) + 1, dtype=np.int64)
        if mats:
            np.cumsum([m.size for m in mats], out=goff[1:])
        cat = np.concatenate([m.ravel() for m in mats] + [np.zeros(0, dtype=np.uint8)])
        px = np.ascontiguousarray(pair_x, dtype=np.int64)
        py = np.ascontiguousarray(pair_y, dtype=np.int64)
        scores = np.ascontiguousarray(scores, dtype=np.int64)
        n_pairs = len(px)
        out_score = np.zeros(n_pairs, dtype=np.int64)
        out_off = np.zeros(n_pairs + 1, dtype=np.int64) if want_cigar else None
        out_cig = np.zeros(max(int(cigar_cap or 32 * n_pairs + 64), 1), dtype=CIGAR_DTYPE) if want_cigar else None
        rc = self._L.gnx_multi_affine_chunk_batch(self._h, _addr(cat), _addr(goff), _addr(nseq), len(mats), _addr(px),
                                                  _addr(py), n_pairs, _addr(scores), int(scores.shape[0]),
                                                  int(gap_open), int(gap_extend), int(chunk), int(bool(want_cigar)),
                                                  _addr(out_score), _addr(out_cig), _addr(out_off),
                                                  0 if out_cig is None else len(out_cig))
        if rc == GNX_ECAP:
            out_cig = np.zeros(max(int(out_off[-1]), 1), dtype=CIGAR_DTYPE)
            rc = self._L.gnx_copy_last_cigars(self._h, _addr(out_cig), len(out_cig))
        self._check(rc)
        if want_cigar:
            return out_score, out_off, out_cig[:int(out_off[-1])]
        return out_score, None, None

    # ---- device-resident entry point (raw device addresses, e.g. torch tensor .data_ptr()) ----
    def batch_device(self, kind, d_alpha_cat, d_alpha_off, d_beta_cat, d_beta_off, alpha_off_host, beta_off_host,
                     n_pairs, scores, gap_open, gap_extend, want_cigar, d_out_score, d_out_cigar=0, d_out_cigar_off=0,
                     cigar_cap=0, d_status=0, stream=0):
        scores = np.ascontiguousarray(scores, dtype=np.int64)
        aoh = None if alpha_off_host is None else np.ascontiguousarray(alpha_off_host, dtype=np.int64)
        boh = None if beta_off_host is None else np.ascontiguousarray(beta_off_host, dtype=np.int64)
        rc = self._L.gnx_batch_device(self._h, int(kind), d_alpha_cat, d_alpha_off, d_beta_cat, d_beta_off,
                                      _addr(aoh), _addr(boh), int(n_pairs), _addr(scores), int(scores.shape[0]),
                                      int(gap_open), int(gap_extend), int(bool(want_cigar)), d_out_score,
                                      d_out_cigar or None, d_out_cigar_off or None, int(cigar_cap), d_status or None,
                                      stream or None)
        self._check(rc)

    def batch_device_twobit(self, kind, d_alpha_words, alpha_len, d_beta_words, beta_len, n_pairs, scores, gap_open,
                            gap_extend, want_cigar, d_out_score, d_out_cigar=0, d_out_cigar_off=0, cigar_cap=0, d_status=0,
                            stream=0):
        """gnx_batch_device_twobit: a uniform batch in dnaTwoBit form, resident on the device."""
        scores = np.ascontiguousarray(scores, dtype=np.int64)
        rc = self._L.gnx_batch_device_twobit(self._h, int(kind), d_alpha_words, int(alpha_len), d_beta_words, int(beta_len),
                                             int(n_pairs), _addr(scores), int(scores.shape[0]), int(gap_open),
                                             int(gap_extend), int(bool(want_cigar)), d_out_score, d_out_cigar or None,
                                             d_out_cigar_off or None, int(cigar_cap), d_status or None, stream or None)
        self._check(rc)


_default_ctx: Optional[Context] = None
_default_lock = threading.Lock()


def default_context() -> Context:
    global _default_ctx
    with _default_lock:
        if _default_ctx is None:
            _default_ctx = Context(0)
        return _default_ctx


def _concat(seqs: Sequence[np.ndarray]) -> Tuple[np.ndarray, np.ndarray]:
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    if len(seqs):
        np.cumsum([len(s) for s in seqs], out=off[1:])
    cat = np.concatenate([np.asarray(s, dtype=np.uint8) for s in seqs]) if len(seqs) and off[-1] > 0 \
        else np.zeros(0, dtype=np.uint8)
    return cat, off


def _split(off: np.ndarray, cig: np.ndarray) -> List[List[Cigar]]:
    rl, op = cig["run_length"], cig["op"]
    return [[Cigar(int(rl[k]), int(op[k])) for k in range(int(off[p]), int(off[p + 1]))] for p in range(len(off) - 1)]


def affine_gap_pairs(alphas, betas, scores, gapOpen, gapExtend, free_end_gaps=False, ctx: Optional[Context] = None):
    """List-of-arrays convenience form: returns [(score, [Cigar...])...]."""
    ctx = ctx or default_context()
    ac, ao = _concat(alphas)
    bc, bo = _concat(betas)
    sc, off, cig = ctx.affine_gap_batch(ac, ao, bc, bo, scores, gapOpen, gapExtend, free_end_gaps)
    return list(zip((int(x) for x in sc), _split(off, cig)))


def const_gap_pairs(alphas, betas, scores, gapPen, ctx: Optional[Context] = None):
    ctx = ctx or default_context()
    ac, ao = _concat(alphas)
    bc, bo = _concat(betas)
    sc, off, cig = ctx.const_gap_batch(ac, ao, bc, bo, scores, gapPen)
    return list(zip((int(x) for x in sc), _split(off, cig)))



class MultiContext:
    """gnx_multi: one context per device behind ONE call (include/gnxalign.h "several GPUs behind one call").
    `devices` may repeat a device (two contexts on it); None = every visible device."""

    def __init__(self, devices: Optional[Sequence[int]] = None, workspace_bytes: int = 0):
        self._L = _lib.load()
        if devices is None:
            self._h = self._L.gnx_multi_create(None, 0, int(workspace_bytes))
        else:
            arr = (C.c_int * len(devices))(*[int(d) for d in devices])
            self._h = self._L.gnx_multi_create(C.cast(arr, C.c_void_p), len(devices), int(workspace_bytes))
        if not self._h:
            raise GnxError(_lib.GNX_ECUDA, self._L.gnx_multi_last_error(None).decode())
        self.n_devices = int(self._L.gnx_multi_device_count(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self._L.gnx_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int):
        if rc != GNX_OK:
            raise GnxError(rc, self._L.gnx_multi_last_error(self._h).decode())

    def set_option(self, name: str, value: int):
        for k in range(self.n_devices):
            rc = self._L.gnx_set_option(self._L.gnx_multi_context(self._h, k), name.encode(), int(value))
            if rc != GNX_OK:
                raise GnxError(rc, f"gnx_set_option({name})")

    def shard_bounds(self, alpha_off, beta_off) -> np.ndarray:
        alpha_off = np.ascontiguousarray(alpha_off, dtype=np.int64)
        beta_off = np.ascontiguousarray(beta_off, dtype=np.int64)
        out = np.zeros(self.n_devices + 1, dtype=np.int64)
        self._check(self._L.gnx_multi_shard_bounds(self._h, _addr(alpha_off), _addr(beta_off), len(alpha_off) - 1, _addr(out)))
        return out

    def _batch(self, kind, alpha_cat, alpha_off, beta_cat, beta_off, scores, gap_open, gap_extend, want_cigar,
               cigar_cap=None, out=None):
        alpha_cat = np.ascontiguousarray(alpha_cat, dtype=np.uint8)
        beta_cat = np.ascontiguousarray(beta_cat, dtype=np.uint8)
        alpha_off = np.ascontiguousarray(alpha_off, dtype=np.int64)
        beta_off = np.ascontiguousarray(beta_off, dtype=np.int64)
        scores = np.ascontiguousarray(scores, dtype=np.int64)
        dim = int(scores.shape[0])
        n_pairs = len(alpha_off) - 1
        if out is not None:
            out_score, out_off, out_cig = out
        else:
            out_score = np.zeros(n_pairs, dtype=np.int64)
            out_off = np.zeros(n_pairs + 1, dtype=np.int64) if want_cigar else None
            out_cig = np.zeros(max(int(cigar_cap or 16 * n_pairs + 64), 1), dtype=CIGAR_DTYPE) if want_cigar else None
        cap = 0 if out_cig is None else len(out_cig)
        if kind == 2:
            rc = self._L.gnx_multi_const_batch(self._h, _addr(alpha_cat), _addr(alpha_off), _addr(beta_cat), _addr(beta_off),
                                               n_pairs, _addr(scores), dim, int(gap_open), int(bool(want_cigar)),
                                               _addr(out_score), _addr(out_cig), _addr(out_off), cap)
        else:
            rc = self._L.gnx_multi_affine_batch(self._h, _addr(alpha_cat), _addr(alpha_off), _addr(beta_cat), _addr(beta_off),
                                                n_pairs, _addr(scores), dim, int(gap_open), int(gap_extend),
                                                GNX_FREE_END if kind == 1 else GNX_GLOBAL, int(bool(want_cigar)),
                                                _addr(out_score), _addr(out_cig), _addr(out_off), cap)
        if rc == GNX_ECAP and out is None:
            out_cig = np.zeros(max(int(out_off[-1]), 1), dtype=CIGAR_DTYPE)
            rc = self._L.gnx_multi_copy_last_cigars(self._h, _addr(out_cig), len(out_cig))
        self._check(rc)
        if want_cigar:
            return out_score, out_off, out_cig[:int(out_off[-1])] if out is None else out_cig
        return out_score, None, None

    def affine_gap_batch(self, alpha_cat, alpha_off, beta_cat, beta_off, scores, gap_open, gap_extend,
                         free_end_gaps=False, want_cigar=True, cigar_cap=None, out=None):
        return self._batch(1 if free_end_gaps else 0, alpha_cat, alpha_off, beta_cat, beta_off, scores, gap_open,
                           gap_extend, want_cigar, cigar_cap, out)

    def const_gap_batch(self, alpha_cat, alpha_off, beta_cat, beta_off, scores, gap_pen, want_cigar=True,
                        cigar_cap=None, out=None):
        return self._batch(2, alpha_cat, alpha_off, beta_cat, beta_off, scores, gap_pen, 0, want_cigar, cigar_cap, out)


# ---- the reference's single-pair API -------------------------------------------------------
def AffineGap_highMem(alpha, beta, scores, gapOpen, gapExtend, ctx=None):
    """align.AffineGap_highMem (align/affineGap_highMem.go:99)."""
    return affine_gap_pairs([alpha], [beta], scores, gapOpen, gapExtend, False, ctx)[0]


def AffineGapLocal(target, query, scores, gapOpen, gapExtend, ctx=None):
    """align.AffineGapLocal (align/affineGap_highMem.go:105): free target overhangs."""
    return affine_gap_pairs([target], [query], scores, gapOpen, gapExtend, True, ctx)[0]


def _require_nonempty(alpha, beta, what):
    if len(alpha) == 0 or len(beta) == 0:
        # the reference's low-mem drivers index out of range / never terminate on an empty input
        raise GnxError(_lib.GNX_EEMPTY, f"{what}: empty sequence (undefined in the reference)")


def _require_one_board(alpha, beta, checkersize_i, checkersize_j, multi_board, what):
    """For inputs that fit one checkerboard the low-memory drivers equal the high-memory result (SURVEY.md 8a:
    verified on 4075 random cases).  Past one board the reference's stitching has defects (cigars differ from
    highMem in ~11 % of random cases and can be invalid) that this library does not reproduce: the caller has to
    ask for the high-memory alignment explicitly instead of getting a silently different answer."""
    if (len(alpha) > checkersize_i or len(beta) > checkersize_j) and multi_board != "highmem":
        raise GnxError(_lib.GNX_EARG,
                       f"{what}: {len(alpha)} x {len(beta)} spans more than one {checkersize_i} x {checkersize_j} "
                       "checkerboard; the reference's multi-board stitching is not reproduced -- pass "
                       "multi_board='highmem' (or call the _highMem function) for the high-memory alignment")


def AffineGap_customizeCheckersize(alpha, beta, scores, gapOpen, gapExtend, checkersize_i, checkersize_j, ctx=None,
                                   multi_board=None):
    """align.AffineGap_customizeCheckersize (align/affineGap.go:73).

    The checkerboard is the reference's memory-saving device, not part of the result for inputs that fit one
    board.  Longer inputs are an explicit error unless multi_board='highmem' (see _require_one_board)."""
    _require_nonempty(alpha, beta, "AffineGap_customizeCheckersize")
    _require_one_board(alpha, beta, checkersize_i, checkersize_j, multi_board, "AffineGap_customizeCheckersize")
    return affine_gap_pairs([alpha], [beta], scores, gapOpen, gapExtend, False, ctx)[0]


def AffineGap(alpha, beta, scores, gapOpen, gapExtend, ctx=None, multi_board=None):
    """align.AffineGap (align/affineGap.go:59): checker size 10000 x 10000."""
    return AffineGap_customizeCheckersize(alpha, beta, scores, gapOpen, gapExtend, 10000, 10000, ctx, multi_board)


def ConstGap_highMem(alpha, beta, scores, gapPen, ctx=None):
    """align.ConstGap_highMem (align/constGap_highMem.go:11)."""
    return const_gap_pairs([alpha], [beta], scores, gapPen, ctx)[0]


def ConstGap_customizeCheckersize(alpha, beta, scores, gapPen, checkersize_i, checkersize_j, ctx=None, multi_board=None):
    """align.ConstGap_customizeCheckersize (align/constGap.go:73); multi-board inputs: see _require_one_board."""
    _require_nonempty(alpha, beta, "ConstGap_customizeCheckersize")
    _require_one_board(alpha, beta, checkersize_i, checkersize_j, multi_board, "ConstGap_customizeCheckersize")
    return const_gap_pairs([alpha], [beta], scores, gapPen, ctx)[0]


def ConstGap(alpha, beta, scores, gapPen, ctx=None, multi_board=None):
    """align.ConstGap (align/constGap.go:13)."""
    return ConstGap_customizeCheckersize(alpha, beta, scores, gapPen, 10000, 10000, ctx, multi_board)


def AffineGapChunk(alpha, beta, scores, gapOpen, gapExtend, chunkSize, ctx=None):
    """align.AffineGapChunk (align/affineGap_highMem.go:227)."""
    ctx = ctx or default_context()
    ac, ao = _concat([alpha])
    bc, bo = _concat([beta])
    sc, off, cig = ctx.affine_gap_chunk_batch(ac, ao, bc, bo, scores, gapOpen, gapExtend, chunkSize)
    return int(sc[0]), _split(off, cig)[0]


# ---- profile DP and the progressive multiple alignment (align/multiAlign.go) --------------------
class Fasta(NamedTuple):
    """fasta.Fasta{Name, Seq} as far as the align package uses it."""
    Name: str
    Seq: np.ndarray


Gap = 10  # dna.Gap


def _stack(group: Sequence[Fasta]) -> np.ndarray:
    return np.stack([np.asarray(f.Seq, dtype=np.uint8) for f in group])


def multipleAffineGapChunk(alpha: Sequence[Fasta], beta: Sequence[Fasta], scores, gapOpen, gapExtend, chunkSize,
                           ctx=None):
    """align.multipleAffineGapChunk (align/affineGap_highMem.go:308-353)."""
    ctx = ctx or default_context()
    sc, off, cig = ctx.multi_affine_chunk_batch([_stack(alpha), _stack(beta)], [0], [1], scores, gapOpen, gapExtend,
                                                chunkSize)
    return int(sc[0]), _split(off, cig)[0]


def multipleAffineGap(alpha: Sequence[Fasta], beta: Sequence[Fasta], scores, gapOpen, gapExtend, ctx=None):
    """align.multipleAffineGap (align/affineGap_highMem.go:272-306)."""
    return multipleAffineGapChunk(alpha, beta, scores, gapOpen, gapExtend, 1, ctx)


def nearestGroupsChunk(groups: List[List[Fasta]], scoreMatrix, gapOpen, gapExtend, chunkSize, ctx=None):
    """align.nearestGroupsChunk (align/multiAlign.go:43-57): every x < y group pair is aligned -- here as ONE
    GPU batch -- and the first pair (in the reference's loop order) with the strictly best score wins.
    Returns (bestX, bestY, bestScore, bestRoute)."""
    ctx = ctx or default_context()
    xs = [x for x in range(len(groups) - 1) for _ in range(x + 1, len(groups))]
    ys = [y for x in range(len(groups) - 1) for y in range(x + 1, len(groups))]
    if not xs:
        return 0, 0, -(1 << 63), []  # bestScore = math.MinInt64, nothing compared
    sc, off, cig = ctx.multi_affine_chunk_batch([_stack(g) for g in groups], xs, ys, scoreMatrix, gapOpen, gapExtend,
                                                chunkSize)
    k = int(np.argmax(sc))  # first occurrence of the maximum == the reference's strict '>' scan
    route = [Cigar(int(r), int(o)) for r, o in cig[int(off[k]):int(off[k + 1])]]
    return xs[k], ys[k], int(sc[k]), route


def nearestGroups(groups, scoreMatrix, gapOpen, gapExtend, ctx=None):
    """align.nearestGroups (align/multiAlign.go:27-41)."""
    return nearestGroupsChunk(groups, scoreMatrix, gapOpen, gapExtend, 1, ctx)


def mergeMultipleAlignments(alpha: Sequence[Fasta], beta: Sequence[Fasta], route) -> List[Fasta]:
    """align.mergeMultipleAlignments (align/multiAlign.go:112-153): host-side column gather."""
    ops = np.repeat(np.array([c[1] for c in route], dtype=np.int64), np.array([c[0] for c in route], dtype=np.int64))
    a_idx = np.cumsum(ops != ColI) - 1  # alpha column consumed by M and D
    b_idx = np.cumsum(ops != ColD) - 1  # beta column consumed by M and I
    out = []
    for f in alpha:
        seq = np.full(len(ops), Gap, dtype=np.uint8)
        seq[ops != ColI] = np.asarray(f.Seq, dtype=np.uint8)[a_idx[ops != ColI]]
        out.append(Fasta(f.Name, seq))
    for f in beta:
        seq = np.full(len(ops), Gap, dtype=np.uint8)
        seq[ops != ColD] = np.asarray(f.Seq, dtype=np.uint8)[b_idx[ops != ColD]]
        out.append(Fasta(f.Name, seq))
    return out


def AllSeqAffineChunk(records: Sequence[Fasta], scoreMatrix, gapOpen, gapExtend, chunkSize, ctx=None) -> List[Fasta]:
    """align.AllSeqAffineChunk (align/multiAlign.go:70-78) incl. mergeFastaGroups' swap-with-last (:20-25)."""
    groups = [[Fasta(r[0], np.asarray(r[1], dtype=np.uint8))] for r in records]
    while len(groups) > 1:
        x, y, _, route = nearestGroupsChunk(groups, scoreMatrix, gapOpen, gapExtend, chunkSize, ctx)
        groups[x] = mergeMultipleAlignments(groups[x], groups[y], route)
        groups[y] = groups[-1]
        groups.pop()
    return groups[0]


def AllSeqAffine(records, scoreMatrix, gapOpen, gapExtend, ctx=None) -> List[Fasta]:
    """align.AllSeqAffine (align/multiAlign.go:59-66)."""
    return AllSeqAffineChunk(records, scoreMatrix, gapOpen, gapExtend, 1, ctx)


# ---- pretty printers (align/view.go) -------------------------------------------------------
def PrintCigar(operations) -> str:
    """align.PrintCigar (align/view.go:25-33)."""
    return "".join(f"{c[0]}{'MID'[c[1]]}" for c in operations)


_RUNES = "ACGTNacgtn-.*"


def View(alpha, beta, operations) -> str:
    """align.View (align/view.go:37-63)."""
    one, two, i, j = [], [], 0, 0
    for run, op in operations:
        for _ in range(run):
            if op == ColM:
                one.append(_RUNES[int(alpha[i])]); two.append(_RUNES[int(beta[j])]); i += 1; j += 1
            elif op == ColI:
                one.append("-"); two.append(_RUNES[int(beta[j])]); j += 1
            else:
                one.append(_RUNES[int(alpha[i])]); two.append("-"); i += 1
    return "".join(one) + "\n" + "".join(two) + "\n"


# ---- streaming engine ------------------------------------------------------------------------
@dataclass
class TargetQueryPair:
    """align.TargetQueryPair (align/affineGap_highMem.go:110-115)."""
    Target: np.ndarray
    Query: np.ndarray
    Score: int = 0
    Cigar: List[Cigar] = field(default_factory=list)


_CLOSE = object()


def GoAffineGapLocalEngine(scores, gapOpen, gapExtend, max_batch: int = 1 << 16, device: int = 0):
    """align.GoAffineGapLocalEngine (align/affineGap_highMem.go:120-179).

    Returns (inputs, outputs) queues (the Go channels, capacity 1000).  A worker thread drains
    whatever is queued (up to max_batch pairs), aligns it as ONE GPU batch with AffineGapLocal
    semantics and emits results in input order (FIFO, as the single reference goroutine does).
    Put `engine_close` (or call inputs.close()) to end the stream; outputs then yields None.  If a batch fails
    (e.g. GNX_EBASE: the reference goroutine would panic), the exception object is put on `outputs` in place of
    that batch's results, the stream is closed (None) and the worker ends -- consumers never block forever."""
    inputs: "queue.Queue" = queue.Queue(maxsize=1000)
    outputs: "queue.Queue" = queue.Queue(maxsize=1000)

    def worker():
        ctx = Context(device)
        try:
            done = False
            while not done:
                item = inputs.get()
                batch = []
                while True:
                    if item is _CLOSE:
                        done = True
                        break
                    batch.append(item)
                    if len(batch) >= max_batch:
                        break
                    try:
                        item = inputs.get_nowait()
                    except queue.Empty:
                        break
                if batch:
                    res = affine_gap_pairs([p.Target for p in batch], [p.Query for p in batch], scores, gapOpen,
                                           gapExtend, True, ctx)
                    for p, (s, c) in zip(batch, res):
                        p.Score, p.Cigar = s, c
                        outputs.put(p)
        except BaseException as exc:  # noqa: BLE001 -- forwarded to the consumer, never swallowed
            outputs.put(exc)
        finally:
            outputs.put(None)  # close(outputs), on success and on failure alike
            ctx.close()

    threading.Thread(target=worker, daemon=True).start()
    inputs.close = lambda: inputs.put(_CLOSE)  # type: ignore[attr-defined]
    return inputs, outputs
