"""Synthetic workload generator -- C, multi-threaded, deterministic per (seed, pair)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libgnxsynth.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "gnx_synth.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-pthread", "-o", _SO, src], check=True)
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.gnx_synth_pairs.argtypes = [C.c_uint64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p,
                                         C.c_void_p, C.c_int]
        _lib.gnx_synth_pairs.restype = C.c_int
        _lib.gnx_pack_uniform.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int]
        _lib.gnx_pack_uniform.restype = C.c_int
    return _lib


def pack_uniform(bases: np.ndarray, n_seqs: int, length: int, out=None, n_threads: int = 0) -> np.ndarray:
    """dnaTwoBit.NewTwoBit of n_seqs sequences of `length` bases (0..3), tightly packed uint64 words (host utility for
    benchmarks / tests: the synthetic batch in the reference's packed form)."""
    wps = (length + 31) // 32
    words = np.empty(n_seqs * wps, dtype=np.uint64) if out is None else out
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    _load().gnx_pack_uniform(bases.ctypes.data, n_seqs, length, words.ctypes.data, n_threads or min(os.cpu_count() or 1, 64))
    return words


def synth_pairs(seed: int, n_pairs: int, n_len: int, m_len: int, first_pair: int = 0, n_threads: int = 0,
                alpha_out=None, beta_out=None):
    """Return (alpha_cat, alpha_off, beta_cat, beta_off) for n_pairs pairs of lengths n_len x m_len.

    alpha_out / beta_out: optional preallocated uint8 buffers (e.g. pinned) of n_pairs*len bytes."""
    n_threads = n_threads or min(os.cpu_count() or 1, 64)
    a = np.empty(n_pairs * n_len, dtype=np.uint8) if alpha_out is None else alpha_out
    b = np.empty(n_pairs * m_len, dtype=np.uint8) if beta_out is None else beta_out
    _load().gnx_synth_pairs(seed, first_pair, n_pairs, n_len, m_len, a.ctypes.data, b.ctypes.data, n_threads)
    ao = np.arange(n_pairs + 1, dtype=np.int64) * n_len
    bo = np.arange(n_pairs + 1, dtype=np.int64) * m_len
    return a, ao, b, bo
