"""gonomics_b200 -- B200 (sm_100a) implementation of gonomics' pairwise-alignment hot path.

`gonomics_b200.align` mirrors the reference's `align` package API over the C ABI of
libgnxalign.so (include/gnxalign.h); `gonomics_b200.genomegraph` and `gonomics_b200.dnatwobit` mirror the gsw
extend / seed step and the 2-bit encoding.  There is no CPU fallback: importing `align` works anywhere,
but every alignment call needs the built library and a CUDA device.
"""
from . import _lib  # noqa: F401
from . import align  # noqa: F401

__all__ = ["align"]
