"""Mirror of the reference package compressor/lz (lzss.go) on top of the C ABI.

Same names, argument meaning and error behaviour as the Go functions a cgo shim would keep:
  CompressAsync(fileContents, useProgressBar, maxSearchBufferLength)   lzss.go:109
  Compress(fileContents, useProgressBar, maxSearchBufferLength)        lzss.go:224
  Decompress(fileContents, useProgressBar)                             lzss.go:323
  NewWriter(w) / NewReader(r)                                          lzss.go:37, 98
Errors are raised as RaisinPanic (the reference panics).
"""
from __future__ import annotations

from . import _lib
from ._lib import RSN_LZSS_ASYNC, RSN_LZSS_ITER

DefaultWindowSize = 4096  # lzss.go:35
Opening, Closing, Separator = "<", ">", ","  # lzss.go:14-18


def CompressAsync(fileContents, useProgressBar: bool = False, maxSearchBufferLength: int = DefaultWindowSize) -> bytes:
    return _lib.call_host(_lib.lib().rsn_lzss_compress, fileContents, int(maxSearchBufferLength), RSN_LZSS_ASYNC)


def Compress(fileContents, useProgressBar: bool = False, maxSearchBufferLength: int = DefaultWindowSize) -> bytes:
    return _lib.call_host(_lib.lib().rsn_lzss_compress, fileContents, int(maxSearchBufferLength), RSN_LZSS_ITER)


def CompressAsyncSharded(fileContents, ngpus: int, maxSearchBufferLength: int = DefaultWindowSize) -> bytes:
    """CompressAsync of one stream with the match search sharded by position range over `ngpus`
    shards (BASELINE configs[4]); byte-identical to CompressAsync."""
    return _lib.call_host(_lib.lib().rsn_lzss_compress_sharded, fileContents, int(maxSearchBufferLength), RSN_LZSS_ASYNC,
                          int(ngpus))


def Decompress(fileContents, useProgressBar: bool = False) -> bytes:
    return _lib.call_host(_lib.lib().rsn_lzss_decompress, fileContents)


class Writer:
    """lz.Writer (lzss.go:29-61): one Write == one CompressAsync of the whole buffer."""

    def __init__(self, w, windowSize: int = DefaultWindowSize):
        self.w = w
        self.windowSize = windowSize
        self.useProgressBar = True

    def Write(self, data) -> int:
        compressed = CompressAsync(data, self.useProgressBar, self.windowSize)
        self.w.write(compressed)
        return len(compressed)

    write = Write

    def Close(self) -> None:
        return None

    close = Close


def NewWriter(w) -> Writer:
    return Writer(w, DefaultWindowSize)


def NewWriterLevel(w, level: int) -> Writer:
    if level < 0:
        raise ValueError(f"lzss: invalid compression level: {level}")
    return Writer(w, level)


class Reader:
    """lz.Reader (lzss.go:63-96): ReadAll of the source, one Decompress, then sliced reads."""

    def __init__(self, r):
        self.r = r
        self.decompressed = None
        self.pos = 0

    def Read(self, size: int = -1) -> bytes:
        if self.decompressed is None:
            self.decompressed = Decompress(self.r.read(), True)
        if size is None or size < 0:
            size = len(self.decompressed) - self.pos
        out = self.decompressed[self.pos:self.pos + size]
        self.pos += len(out)
        return out

    read = Read


def NewReader(r) -> Reader:
    return Reader(r)
