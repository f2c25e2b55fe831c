"""ctypes loader for libgnxalign.so (the C ABI declared in include/gnxalign.h).

The product path never falls back to a CPU implementation: if the library is missing or there
is no CUDA device, loading / context creation raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GNX_LIB") or os.path.join(_HERE, "libgnxalign.so")  # GNX_LIB: A/B builds (tools/)

GNX_OK, GNX_EBASE, GNX_ECAP, GNX_ECHUNK, GNX_EEMPTY, GNX_ECUDA, GNX_EARG, GNX_ERANGE, GNX_EDIVZERO, GNX_EOFFSET, GNX_EINDEX = range(11)
GNX_GLOBAL, GNX_FREE_END = 0, 1
GNX_EXT_LEFT, GNX_EXT_RIGHT, GNX_EXT_LEFT_LOCAL, GNX_EXT_RIGHT_LOCAL = 1, 2, 3, 4
GNX_MATCH_RIGHT, GNX_MATCH_LEFT = 0, 1

# every symbol include/gnxalign.h declares (tests check that the library exports all of them)
EXPORTS = [
    "gnx_device_count", "gnx_create", "gnx_destroy", "gnx_last_error", "gnx_version", "gnx_host_alloc",
    "gnx_host_free", "gnx_affine_batch", "gnx_affine_batch_twobit", "gnx_const_batch", "gnx_affine_chunk_batch", "gnx_copy_last_cigars",
    "gnx_multi_affine_chunk_batch", "gnx_extend_batch", "gnx_batch_device", "gnx_batch_device_twobit", "gnx_launch_count", "gnx_last_fill_stats", "gnx_last_kernel_path", "gnx_set_option",
    "gnx_twobit_new", "gnx_twobit_free", "gnx_twobit_info", "gnx_twobit_download", "gnx_twobit_unpack", "gnx_twobit_get_bases",
    "gnx_twobit_count_matches", "gnx_twobit_pack_device", "gnx_pack_twobit_host", "gnx_seed_index_new", "gnx_seed_index_free", "gnx_seed_index_info",
    "gnx_seed_index_download", "gnx_seed_batch", "gnx_gsw_batch",
    "gnx_multi_create", "gnx_multi_destroy", "gnx_multi_device_count", "gnx_multi_last_error", "gnx_multi_context",
    "gnx_multi_shard_bounds", "gnx_multi_affine_batch", "gnx_multi_const_batch", "gnx_multi_copy_last_cigars",
]


class GnxCigar(C.Structure):
    """align.Cigar{RunLength int64; Op ColType} (align/align.go:21-24)."""
    _fields_ = [("run_length", C.c_int64), ("op", C.c_uint8)]


CIGAR_DTYPE = np.dtype({"names": ["run_length", "op"], "formats": ["<i8", "u1"], "offsets": [0, 8], "itemsize": 16})
# gnx_seed: genomeGraph.SeedDev without NextPart (genomeGraph/index.go:11-19)
SEED_DTYPE = np.dtype([("target_id", "<u4"), ("target_start", "<u4"), ("query_start", "<u4"), ("length", "<u4"),
                       ("pos_strand", "<u4"), ("total_length", "<u4")])

# gnx_giraf: the part of giraf.Giraf the aligner computes (include/gnxalign.h)
GIRAF_DTYPE = np.dtype([("q_start", "<i4"), ("q_end", "<i4"), ("pos_strand", "<i4"), ("t_start", "<i4"), ("t_end", "<i4"),
                        ("node", "<i4"), ("aln_score", "<i8"), ("flag", "<i4"), ("n_cigar", "<i4"), ("cigar_off", "<i8")])

_lib = None


class GnxError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"gnxalign error {code}: {msg}")
        self.code = code


def load() -> C.CDLL:
    """dlopen libgnxalign.so; raise if it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m gonomics_b200.build` (nvcc, sm_100a). "
            "gonomics_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    u8p, i64p, cgp, vp = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p  # raw addresses (host or device)
    i64, ci = C.c_int64, C.c_int
    L.gnx_device_count.restype = ci
    L.gnx_create.argtypes = [ci, C.c_size_t]
    L.gnx_create.restype = vp
    L.gnx_destroy.argtypes = [vp]
    L.gnx_destroy.restype = None
    L.gnx_last_error.argtypes = [vp]
    L.gnx_last_error.restype = C.c_char_p
    L.gnx_version.restype = C.c_char_p
    L.gnx_host_alloc.argtypes = [C.c_size_t]
    L.gnx_host_alloc.restype = vp
    L.gnx_host_free.argtypes = [vp]
    L.gnx_host_free.restype = None
    L.gnx_affine_batch.argtypes = [vp, u8p, i64p, u8p, i64p, i64, i64p, ci, i64, i64, ci, ci, i64p, cgp, i64p, i64]
    L.gnx_affine_batch.restype = ci
    L.gnx_const_batch.argtypes = [vp, u8p, i64p, u8p, i64p, i64, i64p, ci, i64, ci, i64p, cgp, i64p, i64]
    L.gnx_const_batch.restype = ci
    L.gnx_affine_batch_twobit.argtypes = [vp, vp, i64p, i64, vp, i64p, i64, i64, i64p, ci, i64, i64, ci, ci, i64p, cgp, i64p, i64]
    L.gnx_affine_batch_twobit.restype = ci
    L.gnx_batch_device_twobit.argtypes = [vp, ci, vp, i64, vp, i64, i64, i64p, ci, i64, i64, ci, vp, vp, vp, i64, vp, vp]
    L.gnx_batch_device_twobit.restype = ci
    L.gnx_affine_chunk_batch.argtypes = [vp, u8p, i64p, u8p, i64p, i64, i64p, ci, i64, i64, i64, i64p, cgp, i64p, i64]
    L.gnx_affine_chunk_batch.restype = ci
    L.gnx_multi_affine_chunk_batch.argtypes = [vp, u8p, i64p, i64p, i64, i64p, i64p, i64, i64p, ci, i64, i64, i64, ci,
                                               i64p, cgp, i64p, i64]
    L.gnx_multi_affine_chunk_batch.restype = ci
    L.gnx_extend_batch.argtypes = [vp, ci, u8p, i64p, u8p, i64p, i64, i64p, ci, i64, ci, i64p, i64p, i64p, cgp, i64p, i64]
    L.gnx_extend_batch.restype = ci
    L.gnx_copy_last_cigars.argtypes = [vp, cgp, i64]
    L.gnx_copy_last_cigars.restype = ci
    L.gnx_batch_device.argtypes = [vp, ci, u8p, i64p, u8p, i64p, i64p, i64p, i64, i64p, ci, i64, i64, ci, i64p, cgp,
                                   i64p, i64, vp, vp]
    L.gnx_batch_device.restype = ci
    L.gnx_launch_count.argtypes = [vp]
    L.gnx_launch_count.restype = i64
    L.gnx_last_fill_stats.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(i64), C.POINTER(i64)]
    L.gnx_last_fill_stats.restype = ci
    L.gnx_set_option.argtypes = [vp, C.c_char_p, i64]
    L.gnx_set_option.restype = ci
    L.gnx_twobit_new.argtypes = [vp, u8p, i64p, i64, ci, C.POINTER(vp)]
    L.gnx_twobit_new.restype = ci
    L.gnx_twobit_free.argtypes = [vp]
    L.gnx_twobit_free.restype = None
    L.gnx_twobit_info.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
    L.gnx_twobit_info.restype = ci
    L.gnx_twobit_download.argtypes = [vp, vp, vp, i64p, i64p]
    L.gnx_twobit_download.restype = ci
    L.gnx_twobit_unpack.argtypes = [vp, vp, u8p, i64]
    L.gnx_twobit_unpack.restype = ci
    L.gnx_twobit_get_bases.argtypes = [vp, vp, i64p, i64p, i64, u8p]
    L.gnx_twobit_get_bases.restype = ci
    L.gnx_twobit_count_matches.argtypes = [vp, ci, vp, vp, i64p, i64p, i64p, i64p, i64, i64p]
    L.gnx_twobit_count_matches.restype = ci
    L.gnx_twobit_pack_device.argtypes = [vp, u8p, i64, ci, vp, vp]
    L.gnx_twobit_pack_device.restype = ci
    L.gnx_pack_twobit_host.argtypes = [u8p, i64, i64, vp]
    L.gnx_pack_twobit_host.restype = ci
    L.gnx_seed_index_new.argtypes = [vp, u8p, i64p, i64, ci, ci, C.POINTER(vp)]
    L.gnx_seed_index_new.restype = ci
    L.gnx_seed_index_free.argtypes = [vp]
    L.gnx_seed_index_free.restype = None
    L.gnx_seed_index_info.argtypes = [vp, C.POINTER(i64)]
    L.gnx_seed_index_info.restype = ci
    L.gnx_seed_index_download.argtypes = [vp, vp, vp, vp]
    L.gnx_seed_index_download.restype = ci
    L.gnx_seed_batch.argtypes = [vp, vp, u8p, i64p, i64, vp, i64p, i64]
    L.gnx_seed_batch.restype = ci
    L.gnx_gsw_batch.argtypes = [vp, vp, u8p, i64p, i64, i64p, ci, ci, vp, cgp, i64, vp]
    L.gnx_gsw_batch.restype = ci
    L.gnx_last_kernel_path.argtypes = [vp, vp, vp]
    L.gnx_last_kernel_path.restype = ci
    L.gnx_multi_create.argtypes = [vp, ci, C.c_size_t]
    L.gnx_multi_create.restype = vp
    L.gnx_multi_destroy.argtypes = [vp]
    L.gnx_multi_destroy.restype = None
    L.gnx_multi_device_count.argtypes = [vp]
    L.gnx_multi_device_count.restype = ci
    L.gnx_multi_last_error.argtypes = [vp]
    L.gnx_multi_last_error.restype = C.c_char_p
    L.gnx_multi_context.argtypes = [vp, ci]
    L.gnx_multi_context.restype = vp
    L.gnx_multi_shard_bounds.argtypes = [vp, i64p, i64p, i64, i64p]
    L.gnx_multi_shard_bounds.restype = ci
    L.gnx_multi_affine_batch.argtypes = [vp, u8p, i64p, u8p, i64p, i64, i64p, ci, i64, i64, ci, ci, i64p, cgp, i64p, i64]
    L.gnx_multi_affine_batch.restype = ci
    L.gnx_multi_const_batch.argtypes = [vp, u8p, i64p, u8p, i64p, i64, i64p, ci, i64, ci, i64p, cgp, i64p, i64]
    L.gnx_multi_const_batch.restype = ci
    L.gnx_multi_copy_last_cigars.argtypes = [vp, cgp, i64]
    L.gnx_multi_copy_last_cigars.restype = ci
    _lib = L
    return L
