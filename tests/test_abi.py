"""The C-ABI library loads and exports every symbol include/surfface_b200.h declares (CPU only:
no compute call is made)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "surfface_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sfb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_surface():
    names = declared_symbols()
    for must in ("sfb_ctx_create", "sfb_knn_build", "sfb_adjacency_build", "sfb_sparsify_sfgrass",
                 "sfb_laplacian_build", "sfb_lambda", "sfb_diffuse", "sfb_build_laplacian_matrix",
                 "sfb_compute_taumode_lambdas", "sfb_knn_allgather", "sfb_lambda_allgather"):
        assert must in names


def test_library_exports_every_declared_symbol(sfb):
    _ffi = sfb._ffi
    L = _ffi.lib()
    names = declared_symbols()
    assert names, "no declarations parsed"
    for n in names:
        assert hasattr(L, n), f"{n} declared in the header but not exported"
        assert n in _ffi.SYMBOLS, f"{n} has no ctypes signature"
    assert set(_ffi.SYMBOLS) == set(names)
    assert L.sfb_abi_version() == 4


def test_no_cpu_fallback(sfb):
    """Without a CUDA device the product refuses to run instead of computing on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(sfb.SfbError):
        sfb.Context(0)


def test_product_does_not_reference_oracle():
    """Nothing under the product package or include/ may import, include or link the oracle."""
    bad = []
    for base in ("matternet-rs_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".rs", "Makefile", ".toml")):
                    text = open(os.path.join(dirpath, f), errors="ignore").read()
                    if re.search(r"liboracle|import\s+oracle|from\s+oracle|oracle\.c|orc_[a-z]", text):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_rust_sys_crate_declares_every_symbol():
    """The shipped-as-source Rust -sys crate (no Rust toolchain here) stays in sync with the header."""
    rs = open(os.path.join(ROOT, "matternet-rs_b200", "rust", "surfface-b200-sys", "src", "lib.rs")).read()
    declared = set(re.findall(r"pub fn (sfb_[a-z0-9_]+)\(", rs))
    assert declared == set(declared_symbols())

def test_header_is_plain_c_and_links(tmp_path):
    """The boundary is a C ABI: the header compiles as strict C99 and a C program links against the library and runs the
    calls that need no device (the version, the error text of a refused context)."""
    import subprocess
    hdr_dir = os.path.join(ROOT, "include")
    lib_dir = os.path.join(ROOT, "matternet-rs_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c",
                           os.path.join(hdr_dir, "surfface_b200.h")])
    src = tmp_path / "abi_smoke.c"
    src.write_text(
        '#include <stdio.h>\n#include "surfface_b200.h"\n'
        "int main(void) {\n"
        "    sfb_knn_params p; sfb_knn_stats st; sfb_stage_times tm;\n"
        "    (void)p; (void)st; (void)tm;\n"
        '    printf("%d %d\\n", (int)sfb_abi_version(), SFB_ABI_VERSION);\n'
        "    return sfb_abi_version() == SFB_ABI_VERSION ? 0 : 1;\n"
        "}\n")
    exe = tmp_path / "abi_smoke"
    subprocess.check_call(["gcc", "-std=c99", "-I", hdr_dir, str(src), "-o", str(exe), "-L", lib_dir, "-lsurfface_b200",
                           "-Wl,-rpath," + lib_dir])
    out = subprocess.check_output([str(exe)]).decode().split()
    assert out[0] == out[1] == "4"
