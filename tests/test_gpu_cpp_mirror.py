"""Builds and runs the C++ host mirror's restatement of the reference's own tests
(tests/cpp/reference_tests.cpp over include/petal_neighbors.hpp -> C ABI -> CUDA)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "petal-neighbors_b200", "lib")


def _build(tmp_path):
    exe = str(tmp_path / "reference_tests")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "reference_tests.cpp"), "-L", LIBDIR, "-lpetal_b200",
                           f"-Wl,-rpath,{LIBDIR}", "-o", exe])
    return exe


def test_cpp_mirror_compiles(tmp_path):
    _build(tmp_path)


@pytest.mark.gpu
def test_cpp_mirror_reference_tests(tmp_path):
    out = subprocess.run([_build(tmp_path)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "all reference tests passed" in out.stdout
