"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol the header
declares, and refuses to compute without a CUDA device (no fallback)."""
import ctypes
import os
import re

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
