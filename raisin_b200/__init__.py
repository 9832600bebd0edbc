"""raisin_b200 — B200 (sm_100a) implementation of go-compression/raisin's LZSS + Huffman
hot path, behind the reference's own lz / huffman / engine interfaces.

The package is a thin host-side mirror (ctypes) of the C ABI in include/raisin_b200.h; all
codec work happens in hand-written CUDA kernels inside libraisin_b200.so.  No CPU fallback.
"""
from . import _lib, engine, huffman, lz  # noqa: F401
from ._lib import RaisinPanic  # noqa: F401

__all__ = ["lz", "huffman", "engine", "RaisinPanic"]
