"""CPU tests of the boundary: the C-ABI library loads and exports every symbol the header
declares, and the product has no CPU fallback (codec calls fail loudly without a device)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported(rsn):
    hdr = open(os.path.join(ROOT, "include", "raisin_b200.h")).read()
    declared = set(re.findall(r"\b(rsn_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    L = rsn._lib.lib()
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert declared == set(rsn._lib.EXPORTS)


def test_version_and_strerror(rsn):
    L = rsn._lib.lib()
    assert b"sm_100a" in L.rsn_version()
    assert L.rsn_strerror(0) == b"ok"
    for rc in (-1, -2, -3, -4, -5, -10, -11, -12, -13, -14, -15, -16):
        assert L.rsn_strerror(rc) != b"unknown error"


def test_no_cpu_fallback(rsn):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the fail-loudly path is for GPU-less hosts")
    with pytest.raises(rsn.RaisinPanic) as ei:
        rsn.lz.CompressAsync(b"hello hello hello")
    assert ei.value.rc in (-1, -2)
    with pytest.raises(rsn.RaisinPanic):
        rsn.huffman.Compress(b"hello")


def test_product_does_not_import_oracle():
    """Nothing under raisin_b200/ may reference oracle/."""
    pkg = os.path.join(ROOT, "raisin_b200")
    for dirpath, _, files in os.walk(pkg):
        if "_obj" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "pyoracle" not in src and "raisin_oracle" not in src and "rsno_" not in src, f


def test_invalid_args(rsn):
    import ctypes as C

    L = rsn._lib.lib()
    out = C.POINTER(C.c_uint8)()
    n = C.c_size_t(0)
    assert L.rsn_lzss_compress(None, 5, 4096, 0, C.byref(out), C.byref(n)) == -4
    assert L.rsn_compress_layers(b"lzss,zip", b"abc", 3, C.byref(out), C.byref(n)) == -4
