"""Host-side mirror of gonomics' `dna/dnaTwoBit` package over the libgnxalign C ABI.

    TwoBit, NewTwoBit, GetBase            dna/dnaTwoBit/dnaTwoBit.go:14-17,59-78
    NewTwoBitRainbow                      dna/dnaTwoBit/rainbow.go:8-25
    CountRightMatches, CountLeftMatches   dna/dnaTwoBit/perfectAlign.go:10-85

A `TwoBitSet` is a batch of TwoBit sequences resident on the GPU (a genome's nodes, a batch of reads): the
bytes are uploaded and packed once, queries run against the device copy.  `TwoBit` is the reference's
single-sequence type as a view into a set.  Every call goes through the CUDA library; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import GNX_MATCH_LEFT, GNX_MATCH_RIGHT
from .align import Context, _concat, default_context

A, C_, G, T = 0, 1, 2, 3  # dnaTwoBit.go:19-24


class TwoBitSet:
    """n sequences packed by NewTwoBit (lead = 0) or as element `lead` of NewTwoBitRainbow, on the device."""

    def __init__(self, seq_cat: np.ndarray, seq_off: np.ndarray, lead: int = 0, ctx: Optional[Context] = None):
        self.ctx = ctx or default_context()
        self._L = self.ctx._L
        cat = np.ascontiguousarray(seq_cat, dtype=np.uint8)
        off = np.ascontiguousarray(seq_off, dtype=np.int64)
        h = C.c_void_p(None)
        self._h = None
        self.ctx._check(self._L.gnx_twobit_new(self.ctx._h, cat.ctypes.data, off.ctypes.data, len(off) - 1, int(lead),
                                               C.byref(h)))
        self._h = h
        n, w = C.c_int64(0), C.c_int64(0)
        self._L.gnx_twobit_info(self._h, C.byref(n), C.byref(w))
        self.n_seqs, self.total_words = n.value, w.value

    @classmethod
    def from_seqs(cls, seqs: Sequence[np.ndarray], lead: int = 0, ctx: Optional[Context] = None) -> "TwoBitSet":
        cat, off = _concat(seqs)
        return cls(cat, off, lead, ctx)

    def close(self):
        if getattr(self, "_h", None) and getattr(self.ctx, "_h", None):
            self._L.gnx_twobit_free(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def download(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """(Seq words of all sequences concatenated, word offsets [n+1], Len [n])."""
        words = np.zeros(max(self.total_words, 1), dtype=np.uint64)
        woff = np.zeros(self.n_seqs + 1, dtype=np.int64)
        ln = np.zeros(max(self.n_seqs, 1), dtype=np.int64)
        self.ctx._check(self._L.gnx_twobit_download(self.ctx._h, self._h, words.ctypes.data, woff.ctypes.data, ln.ctypes.data))
        return words[:self.total_words], woff, ln[:self.n_seqs]

    def unpack(self) -> Tuple[np.ndarray, np.ndarray]:
        """GetBase for every position: (bases concatenated, byte offsets [n+1])."""
        _, _, ln = self.download()
        off = np.zeros(self.n_seqs + 1, dtype=np.int64)
        np.cumsum(ln, out=off[1:])
        out = np.zeros(max(int(off[-1]), 1), dtype=np.uint8)
        self.ctx._check(self._L.gnx_twobit_unpack(self.ctx._h, self._h, out.ctypes.data, int(off[-1])))
        return out[:int(off[-1])], off

    def get_bases(self, q_seq, q_pos) -> np.ndarray:
        qs = np.ascontiguousarray(q_seq, dtype=np.int64)
        qp = np.ascontiguousarray(q_pos, dtype=np.int64)
        out = np.zeros(max(len(qs), 1), dtype=np.uint8)
        self.ctx._check(self._L.gnx_twobit_get_bases(self.ctx._h, self._h, qs.ctypes.data, qp.ctypes.data, len(qs), out.ctypes.data))
        return out[:len(qs)]


def count_matches(direction: int, one: TwoBitSet, two: TwoBitSet, q_one, q_start_one, q_two, q_start_two) -> np.ndarray:
    """Batched CountRightMatches (GNX_MATCH_RIGHT) / CountLeftMatches (GNX_MATCH_LEFT)."""
    arrs = [np.ascontiguousarray(a, dtype=np.int64) for a in (q_one, q_start_one, q_two, q_start_two)]
    n = len(arrs[0])
    out = np.zeros(max(n, 1), dtype=np.int64)
    one.ctx._check(one._L.gnx_twobit_count_matches(one.ctx._h, int(direction), one._h, two._h, *[a.ctypes.data for a in arrs],
                                                   n, out.ctypes.data))
    return out[:n]


class TwoBit:
    """dnaTwoBit.TwoBit{Seq []uint64; Len int}: sequence `idx` of a TwoBitSet."""

    def __init__(self, owner: TwoBitSet, idx: int = 0):
        self.set, self.idx = owner, idx
        self._cache = None

    def _host(self):
        if self._cache is None:
            words, woff, ln = self.set.download()
            self._cache = (words[woff[self.idx]:woff[self.idx + 1]].copy(), int(ln[self.idx]))
        return self._cache

    @property
    def Seq(self) -> np.ndarray:
        return self._host()[0]

    @property
    def Len(self) -> int:
        return self._host()[1]


def pack_uniform_host(bases, count: int, length: int) -> np.ndarray:
    """dnaTwoBit.NewTwoBit of `count` sequences of `length` bases each (back to back in `bases`) on the host threads of
    the library (gnx_pack_twobit_host; no GPU involved): (length + 31) // 32 uint64 words per sequence -- the input
    format of affine_gap_batch_twobit.  Raises GnxError(GNX_EBASE) when a base is >= 4."""
    seq = np.ascontiguousarray(bases, dtype=np.uint8)
    if seq.size < count * length:
        raise ValueError("bases is shorter than count * length")
    words = np.empty(count * ((length + 31) // 32), dtype=np.uint64)
    L = _lib.load()
    rc = L.gnx_pack_twobit_host(seq.ctypes.data_as(L.gnx_pack_twobit_host.argtypes[0]), count, length, words.ctypes.data)
    if rc != _lib.GNX_OK:
        raise _lib.GnxError(rc, "gnx_pack_twobit_host: a base is >= 4" if rc == _lib.GNX_EBASE else "gnx_pack_twobit_host failed")
    return words


def NewTwoBit(inSeq, ctx: Optional[Context] = None) -> TwoBit:
    """dnaTwoBit.NewTwoBit (dnaTwoBit.go:68)."""
    return TwoBit(TwoBitSet.from_seqs([np.asarray(inSeq, dtype=np.uint8)], 0, ctx))


def NewTwoBitRainbow(inSeq, ctx: Optional[Context] = None) -> List[TwoBit]:
    """dnaTwoBit.NewTwoBitRainbow (rainbow.go:8): the 32 encodings with 0..31 leading 'A's."""
    seq = np.asarray(inSeq, dtype=np.uint8)
    return [TwoBit(TwoBitSet.from_seqs([seq], k, ctx)) for k in range(32)]


def GetBase(frag: TwoBit, pos: int) -> int:
    """dnaTwoBit.GetBase (dnaTwoBit.go:59)."""
    return int(frag.set.get_bases([frag.idx], [pos])[0])


def CountRightMatches(one: TwoBit, startOne: int, two: TwoBit, startTwo: int) -> int:
    """dnaTwoBit.CountRightMatches (perfectAlign.go:10)."""
    return int(count_matches(GNX_MATCH_RIGHT, one.set, two.set, [one.idx], [startOne], [two.idx], [startTwo])[0])


def CountLeftMatches(one: TwoBit, startOne: int, two: TwoBit, startTwo: int) -> int:
    """dnaTwoBit.CountLeftMatches (perfectAlign.go:49)."""
    return int(count_matches(GNX_MATCH_LEFT, one.set, two.set, [one.idx], [startOne], [two.idx], [startTwo])[0])
