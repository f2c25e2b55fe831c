"""Build libgnxalign.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgnxalign.so")
SOURCES = ["gnx_api.cu", "gnx_pack_host.cpp"]
DEPS = ["gnx_api.cu", "gnx_pack_host.cpp", "gnx_kernels.cuh", "gnx_fill3.cuh", "gnx_fill16.cuh", "gnx_ckpt.cuh", "gnx_long.cuh", "gnx_profile.cuh", "gnx_twobit.cuh", "gnx_twobit_api.inl", "gnx_multi.inl", "gnx_gsw.inl", os.path.join("..", "..", "include", "gnxalign.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--use_fast_math", "-Xcompiler", "-fPIC,-O2,-Wall,-pthread", "-shared", "-cudart", "shared"]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    return "nvcc"


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    extra = os.environ.get("GNX_NVCC_EXTRA", "").split()
    cmd = [nvcc(), *NVCC_FLAGS, *extra, "-o", LIB, *[os.path.join(CSRC, s) for s in SOURCES]]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(LIB)
