"""The C-ABI boundary: header <-> shared library <-> ctypes mirror (CPU only, no compute)."""
import ctypes
import os
import re

import pytest

import halo2_pse_b200 as h
from halo2_pse_b200 import _ffi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "halo2_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(h2b_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_what_ffi_binds():
    assert header_functions() == sorted(_ffi.SYMBOLS)


def test_product_library_exports_every_declared_symbol():
    path = build.build_product()  # nvcc cross-compiles sm_100a without a GPU
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    for name in header_functions():
        assert hasattr(lib, name), name
    _ffi.load(path)


def test_header_compiles_as_c():
    import subprocess
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write('#include "halo2_b200.h"\nint main(void){h2b_fr x; (void)x; return sizeof(h2b_g1)==96?0:1;}\n')
        subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), c, "-o",
                        os.path.join(d, "t")], check=True)
        subprocess.run([os.path.join(d, "t")], check=True)


def test_no_cpu_fallback_without_a_device():
    """On a box without a GPU the product fails loudly instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    build.build_product()
    with pytest.raises(h.H2BError) as e:
        h.Context(0)
    assert e.value.code == h.H2B_ERR_CUDA


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(h.H2BError):
        _ffi.load(str(tmp_path / "nope.so"))


def test_product_does_not_reference_the_oracle():
    """oracle/ is test infrastructure: nothing under the package may import or link it."""
    pkg = os.path.join(ROOT, "halo2-pse_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower() or f == "build.py", os.path.join(dirpath, f)
