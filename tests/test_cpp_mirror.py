"""The C++ host-side mirror (include/surfface_b200.hpp) over the C ABI: it compiles against the header (CPU), and its
test program -- the reference's own tests restated, plus parity with the oracle -- passes on a B200 (GPU)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_mirror.cpp")
BIN = os.path.join(ROOT, "tests", "cpp", "test_mirror")


def build(oracle):
    oracle.build()
    lib_dir = os.path.join(ROOT, "matternet-rs_b200")
    orc_dir = os.path.join(ROOT, "oracle")
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), SRC, "-o", BIN,
           "-L", lib_dir, "-l:libsurfface_b200.so", "-L", orc_dir, "-l:liboracle.so",
           f"-Wl,-rpath,{lib_dir}", f"-Wl,-rpath,{orc_dir}", "-lgomp"]
    subprocess.check_call(cmd)
    return BIN


def test_cpp_mirror_compiles_and_links(oracle, sfb):
    assert os.path.exists(build(oracle))


@pytest.mark.gpu
def test_cpp_mirror_runs(oracle, sfb):
    out = subprocess.run([build(oracle)], capture_output=True, text=True, timeout=300)
    print(out.stdout[-3000:], out.stderr[-2000:])
    assert out.returncode == 0 and "all C++ mirror tests passed" in out.stdout
