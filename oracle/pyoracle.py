"""ctypes binding of oracle/_build/libraisin_oracle.so (the C restatement of the reference).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package raisin_b200 never imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libraisin_oracle.so")

ERRORS = {
    -1: "nomem",
    -2: "empty_input",
    -3: "no_separator",
    -4: "bad_header",
    -5: "truncated",
    -6: "guard",
    -7: "bad_reference",
    -8: "single_leaf_loop",
}


class OracleError(Exception):
    def __init__(self, code: int):
        super().__init__(f"oracle error {code} ({ERRORS.get(code, '?')})")
        self.code = code
        self.name = ERRORS.get(code, "?")


def build(force: bool = False) -> str:
    """Compile the C oracle with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "raisin_oracle.c")
    hdr = os.path.join(_HERE, "raisin_oracle.h")
    stale = (
        force
        or not os.path.exists(_SO)
        or (os.path.exists(src) and os.path.getmtime(_SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)))
    )
    if stale:
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u8p = C.POINTER(C.c_uint8)
        L.rsno_free.argtypes = [C.c_void_p]
        L.rsno_free.restype = None
        for name in ("rsno_escape", "rsno_unescape", "rsno_lzss_decompress", "rsno_huff_compress"):
            f = getattr(L, name)
            f.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(u8p), C.POINTER(C.c_size_t)]
            f.restype = C.c_int
        L.rsno_lzss_compress_async.argtypes = [
            C.c_char_p, C.c_size_t, C.c_int64, C.c_int, C.c_int, C.POINTER(u8p), C.POINTER(C.c_size_t)]
        L.rsno_lzss_compress_async.restype = C.c_int
        L.rsno_lzss_compress_iter.argtypes = [C.c_char_p, C.c_size_t, C.c_int64, C.POINTER(u8p), C.POINTER(C.c_size_t)]
        L.rsno_lzss_compress_iter.restype = C.c_int
        L.rsno_huff_decompress.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.POINTER(u8p), C.POINTER(C.c_size_t)]
        L.rsno_huff_decompress.restype = C.c_int
        L.rsno_lzss_match_arrays.argtypes = [
            C.c_char_p, C.c_size_t, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.rsno_lzss_match_arrays.restype = C.c_int
        L.rsno_utf8_decode.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p]
        L.rsno_utf8_decode.restype = C.c_size_t
        L.rsno_huff_code_table.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        L.rsno_huff_code_table.restype = C.c_int
        _lib = L
    return _lib


def _take(rc, out, out_n) -> bytes:
    if rc != 0:
        raise OracleError(rc)
    try:
        n = out_n.value
        if n < (1 << 31) - 1:
            return C.string_at(out, n)
        # ctypes.string_at takes a C int
        return bytes((C.c_ubyte * n).from_address(C.cast(out, C.c_void_p).value))
    finally:
        lib().rsno_free(out)


def _call(fn, data: bytes, *extra) -> bytes:
    out = C.POINTER(C.c_uint8)()
    out_n = C.c_size_t(0)
    rc = fn(data, len(data), *extra, C.byref(out), C.byref(out_n))
    return _take(rc, out, out_n)


def escape(data: bytes) -> bytes:
    return _call(lib().rsno_escape, data)


def unescape(data: bytes) -> bytes:
    return _call(lib().rsno_unescape, data)


def lzss_compress_async(data: bytes, window: int = 4096, literal: bool = False, threads: int = 1) -> bytes:
    return _call(lib().rsno_lzss_compress_async, data, window, 1 if literal else 0, threads)


def lzss_compress_iter(data: bytes, window: int = 4096) -> bytes:
    return _call(lib().rsno_lzss_compress_iter, data, window)


def lzss_decompress(data: bytes) -> bytes:
    return _call(lib().rsno_lzss_decompress, data)


def huff_compress(data: bytes) -> bytes:
    return _call(lib().rsno_huff_compress, data)


def huff_decompress(data: bytes, strict: bool = False) -> bytes:
    return _call(lib().rsno_huff_decompress, data, 1 if strict else 0)


def lzss_match_arrays(enc: bytes, window: int = 4096, literal: bool = False, threads: int = 1):
    """(len[i], off[i]) of compressorWorker over the already-escaped buffer."""
    n = len(enc)
    ln = np.zeros(n, dtype=np.uint32)
    off = np.zeros(n, dtype=np.uint32)
    rc = lib().rsno_lzss_match_arrays(enc, n, window, 1 if literal else 0, threads, ln.ctypes.data, off.ctypes.data)
    if rc:
        raise OracleError(rc)
    return ln, off


def utf8_decode(data: bytes) -> np.ndarray:
    out = np.zeros(max(len(data), 1), dtype=np.int32)
    m = lib().rsno_utf8_decode(data, len(data), out.ctypes.data)
    return out[:m].copy()


def huff_code_table(runes, freqs):
    runes = np.ascontiguousarray(runes, dtype=np.int32)
    freqs = np.ascontiguousarray(freqs, dtype=np.uint64)
    k = len(runes)
    code = np.zeros(k, dtype=np.uint64)
    ln = np.zeros(k, dtype=np.uint8)
    rc = lib().rsno_huff_code_table(runes.ctypes.data, freqs.ctypes.data, k, code.ctypes.data, ln.ctypes.data)
    if rc:
        raise OracleError(rc)
    return code, ln


# ---- helpers shared by parity tests ------------------------------------------------------


def huff_split(blob: bytes):
    """(header bytes, payload bytes after the 5C 0A separator)."""
    k = blob.find(b"\\\n")
    if k < 0:
        raise ValueError("no separator")
    return blob[:k], blob[k + 2:]


def huff_header_map(header: bytes) -> dict:
    """Parse a header the way decodeTree does (huffman.go:196-227) into {rune: freq}."""
    from . import go_literal

    return go_literal.decodeTree(header)
