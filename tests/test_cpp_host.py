"""The C++ host mirrors (gonomics_b200/csrc/host/align.hpp, genomegraph.hpp): compile on CPU, run on the GPU box."""
import os
import subprocess

import pytest

from gonomics_b200 import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TESTS = ["test_align_host", "test_genomegraph_host"]


def _compile(name):
    build.build()
    libdir = os.path.join(ROOT, "gonomics_b200")
    src = os.path.join(ROOT, "tests", "cpp", name + ".cpp")
    exe = os.path.join(ROOT, "tests", "cpp", name)
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-pthread", src, "-o", exe, f"-L{libdir}", "-lgnxalign",
                    f"-Wl,-rpath,{libdir}", "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"], check=True)
    return exe


@pytest.mark.parametrize("name", TESTS)
def test_cpp_host_mirror_compiles_and_links(name):
    assert os.path.exists(_compile(name))


@pytest.mark.gpu
@pytest.mark.parametrize("name", TESTS)
def test_cpp_host_mirror_replays_reference_tests(name):
    r = subprocess.run([_compile(name)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
