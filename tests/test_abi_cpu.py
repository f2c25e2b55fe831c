"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol the header
declares, and refuses to compute without a CUDA device (no fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from gonomics_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()  # nvcc cross-compiles sm_100a without a GPU
    return _lib.load()


def test_header_symbols_all_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "gnxalign.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(gnx_[a-z_0-9]+)\s*\(", hdr)))
    assert declared == sorted(_lib.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None, name


def test_cigar_layout_matches_go_struct():
    # align.Cigar{RunLength int64; Op uint8}: 16 bytes, Op at offset 8 (amd64)
    assert ctypes.sizeof(_lib.GnxCigar) == 16
    assert _lib.GnxCigar.op.offset == 8
    assert _lib.CIGAR_DTYPE.itemsize == 16


def test_version_and_error_strings(lib):
    assert b"sm_100a" in lib.gnx_version()


def test_no_silent_cpu_fallback(lib):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present: covered by the gpu tests")
    assert lib.gnx_device_count() == 0
    from gonomics_b200 import align
    with pytest.raises(_lib.GnxError):
        align.Context(0)


def test_product_does_not_import_oracle():
    # the oracle is test infrastructure; nothing under gonomics_b200/ may reference it
    pkg = os.path.join(ROOT, "gonomics_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "gnx_oracle" not in txt, f


def test_host_packer_matches_the_oracle():
    """gnx_pack_twobit_host (the packer gnx_affine_batch runs while it stages pageable bytes; pure host code, no GPU):
    word for word dnaTwoBit.NewTwoBit of every sequence (the oracle's packer, itself pinned to the reference's
    known answers in test_twobit_oracle.py), for lengths around the word and 4-base boundaries, many sequences
    (several threads) and the GNX_EBASE report for a base >= 4."""
    import oracle as orc
    from gonomics_b200 import _lib, dnatwobit
    rng = np.random.default_rng(11)
    for length, count in ((1, 5), (3, 7), (4, 3), (31, 9), (32, 9), (33, 9), (64, 4), (150, 1000), (500, 333), (97, 70000)):
        seqs = rng.integers(0, 4, size=(count, length), dtype=np.uint8)
        words = dnatwobit.pack_uniform_host(seqs.reshape(-1), count, length)
        wl = (length + 31) // 32
        assert words.shape == (count * wl,)
        for p in list(range(min(count, 40))) + [count - 1, count // 2]:
            want = orc.new_twobit(seqs[p])[0][:wl]
            assert np.array_equal(words[p * wl:(p + 1) * wl], want), (length, count, p)
    bad = rng.integers(0, 4, size=(50000, 40), dtype=np.uint8)
    bad[49990, 39] = 4
    with pytest.raises(_lib.GnxError) as ei:
        dnatwobit.pack_uniform_host(bad.reshape(-1), 50000, 40)
    assert ei.value.code == _lib.GNX_EBASE
    assert dnatwobit.pack_uniform_host(np.zeros(0, dtype=np.uint8), 0, 10).size == 0


def test_host_packer_simd_equals_scalar(lib):
    """The run-time-selected (AVX2 where the CPU has it) and the scalar form of the host packer give the same words
    and the same verdict on invalid bases."""
    rng = np.random.default_rng(12)
    for fn in (lib.gnx_pack_range_host, lib.gnx_pack_range_host_scalar):
        fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64]
        fn.restype = ctypes.c_bool
    for length, count in ((500, 2000), (150, 3001), (64, 10), (33, 77), (7, 100)):
        seqs = rng.integers(0, 4, size=count * length, dtype=np.uint8)
        wl = (length + 31) // 32
        a, b = np.zeros(count * wl, np.uint64), np.zeros(count * wl, np.uint64)
        assert lib.gnx_pack_range_host(a.ctypes.data, seqs.ctypes.data, count, length, wl)
        assert lib.gnx_pack_range_host_scalar(b.ctypes.data, seqs.ctypes.data, count, length, wl)
        assert np.array_equal(a, b), (length, count)
        for pos in (0, length // 2, count * length - 1):
            bad = seqs.copy()
            bad[pos] = 4 + (pos % 3) * 60
            assert not lib.gnx_pack_range_host(a.ctypes.data, bad.ctypes.data, count, length, wl)
            assert not lib.gnx_pack_range_host_scalar(b.ctypes.data, bad.ctypes.data, count, length, wl)
