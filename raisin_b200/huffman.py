"""Mirror of the reference package compressor/huffman (huffman.go) on top of the C ABI.

  Compress(fileContents)     huffman.go:299
  Decompress(fileContents)   huffman.go:327
  NewWriter(w) / NewReader(r) huffman.go:372, 395

`strict_limits=True` reproduces the reference's 900000-bit recursion guard (huffman.go:132);
the default lifts it.  Unlike the reference there is no package-level state: two
Decompress calls in one process do not concatenate (huffman.go:129 never resets `answer`).
"""
from __future__ import annotations

from . import _lib


def Compress(fileContents) -> bytes:
    return _lib.call_host(_lib.lib().rsn_huff_compress, fileContents)


def Decompress(fileContents, strict_limits: bool = False) -> bytes:
    return _lib.call_host(_lib.lib().rsn_huff_decompress, fileContents, 1 if strict_limits else 0)


class Writer:
    def __init__(self, w):
        self.w = w

    def Write(self, data) -> int:
        compressed = Compress(data)
        self.w.write(compressed)
        return len(compressed)

    write = Write

    def Close(self) -> None:
        return None

    close = Close


def NewWriter(w) -> Writer:
    return Writer(w)


class Reader:
    def __init__(self, r):
        self.r = r
        self.decompressed = None
        self.pos = 0

    def Read(self, size: int = -1) -> bytes:
        if self.decompressed is None:
            self.decompressed = Decompress(self.r.read())
        if size is None or size < 0:
            size = len(self.decompressed) - self.pos
        out = self.decompressed[self.pos:self.pos + size]
        self.pos += len(out)
        return out

    read = Read


def NewReader(r) -> Reader:
    return Reader(r)
