"""The C++ host mirror (gonomics_b200/csrc/host/align.hpp): compiles on CPU, runs on the GPU box."""
import os
import subprocess

import pytest

from gonomics_b200 import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_align_host.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "test_align_host")


def _compile():
    build.build()
    libdir = os.path.join(ROOT, "gonomics_b200")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-pthread", SRC, "-o", EXE, f"-L{libdir}", "-lgnxalign",
                    f"-Wl,-rpath,{libdir}", "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"], check=True)


def test_cpp_host_mirror_compiles_and_links():
    _compile()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_cpp_host_mirror_replays_reference_tests():
    _compile()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
