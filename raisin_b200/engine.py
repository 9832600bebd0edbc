"""Mirror of the reference's engine layer for the lzss/huffman path (engine/engine.go).

  Writers / Readers registry           engine.go:47-57, 100-110
  CompressedFile.Write / .Read         engine.go:113-139, 60-97
  compress / decompress layer loops    engine.go:443-479
  CompressFile / DecompressFile (.rsn) engine.go:157-199
  BenchmarkFile                        engine.go:357-441 (timed region = compress + decompress)
"""
from __future__ import annotations

import io
import math
import time
from dataclasses import dataclass

from . import _lib, huffman, lz

Writers = {"lzss": lz.NewWriter, "huffman": huffman.NewWriter}
Readers = {"lzss": lz.NewReader, "huffman": huffman.NewReader}


class CompressedFile:
    def __init__(self, CompressionEngine: str = "", Compressed: bytes = b"", MaxSearchBufferLength: int = 4096):
        self.CompressionEngine = CompressionEngine
        self.Compressed = Compressed
        self.Decompressed = None
        self.pos = 0
        self.MaxSearchBufferLength = MaxSearchBufferLength  # set but never read, as in engine.go:44

    def Write(self, content: bytes) -> int:
        b = io.BytesIO()
        w = Writers[self.CompressionEngine](b)
        w.Write(content)
        w.Close()
        compressed = b.getvalue()
        self.Compressed = self.Compressed + compressed
        return len(compressed)

    def Read(self, size: int = -1) -> bytes:
        if self.Decompressed is None:
            r = Readers[self.CompressionEngine](io.BytesIO(self.Compressed))
            self.Decompressed = r.read()
        if size is None or size < 0:
            size = len(self.Decompressed) - self.pos
        out = self.Decompressed[self.pos:self.pos + size]
        self.pos += len(out)
        return out


def compress(content: bytes, algorithms) -> bytes:
    """engine.go:443-452 — one call per layer through the io.Writer plumbing."""
    for algorithm in algorithms:
        f = CompressedFile(CompressionEngine=algorithm, MaxSearchBufferLength=4096)
        f.Write(content)
        content = f.Compressed
    return content


def decompress(content: bytes, algorithms) -> bytes:
    """engine.go:454-479 — layers in reverse order."""
    for algorithm in reversed(list(algorithms)):
        f = CompressedFile(CompressionEngine=algorithm, Compressed=content)
        f.Read()
        content = f.Decompressed
    return content


def compress_fused(content, algorithms) -> bytes:
    """Same result as compress(), but the layers are chained on the device in one C-ABI call."""
    return _lib.call_host(lambda p, n, o, on: _lib.lib().rsn_compress_layers(",".join(algorithms).encode(), p, n, o, on),
                          content)


def decompress_fused(content, algorithms) -> bytes:
    return _lib.call_host(
        lambda p, n, o, on: _lib.lib().rsn_decompress_layers(",".join(algorithms).encode(), p, n, o, on), content)


def batch(files, algorithms, compress_: bool = True, workers: int = 0):
    """Independent files through the layer list on a pool of host threads/streams (one C-ABI call).
    Returns a list with one bytes object per file; a failed file (the reference would panic) is None."""
    import ctypes as C

    L = _lib.lib()
    n = len(files)
    keep = [_lib._as_ptr(f) for f in files]
    ins = (C.c_void_p * n)(*[k[0] for k in keep])
    ns = (C.c_size_t * n)(*[k[1] for k in keep])
    outs = (C.c_void_p * n)()
    out_ns = (C.c_size_t * n)()
    sentinel = -1000  # no code the library returns: "the call never reached this file"
    rcs = (C.c_int * n)(*([sentinel] * n))
    rc = L.rsn_batch_layers(",".join(algorithms).encode(), 1 if compress_ else 0, n, ins, ns, outs, out_ns, rcs, workers, 0)
    if rc != 0 and all(r == sentinel for r in rcs):
        raise _lib.RaisinPanic(rc, L.rsn_strerror(rc).decode())  # failed before any per-file work (unknown layer, no device, ...)
    res = []
    for i in range(n):
        if rcs[i] != 0:
            res.append(None)
        else:
            res.append(_lib.bytes_at(outs[i], out_ns[i]))
            L.rsn_free(outs[i])
    return res


def CompressFile(algorithms, path: str, output: str | None = None) -> str:
    """engine.go:157-166: whole file in, `<path>.rsn` out (no framing, no magic)."""
    with open(path, "rb") as fh:
        content = fh.read()
    out = output or path + ".rsn"
    with open(out, "wb") as fh:
        fh.write(compress(content, algorithms))
    return out


def DecompressFile(algorithms, path: str, output: str) -> str:
    with open(path, "rb") as fh:
        content = fh.read()
    with open(output, "wb") as fh:
        fh.write(decompress(content, algorithms))
    return output


@dataclass
class Result:
    CompressionEngine: str
    TimeTaken: str
    Ratio: float
    ActualEntropy: float
    Entropy: float
    Lossless: bool
    Failed: bool
    Seconds: float = 0.0


def _entropy(counts, total) -> float:
    return -sum((c / total) * math.log(c / total) for c in counts if c)


def BenchmarkFile(algorithms, path_or_bytes, fused: bool = False) -> Result:
    """engine.go:357-441.  The timed region is compress + decompress, wall clock."""
    if isinstance(path_or_bytes, (bytes, bytearray)):
        fileContents = bytes(path_or_bytes)
    else:
        with open(path_or_bytes, "rb") as fh:
            fileContents = fh.read()
    name = ",".join(algorithms)
    if fused:  # everything, histograms and the lossless comparison included, in one C-ABI call on the device
        import ctypes as C

        r = _lib.BenchResult()
        ptr, n, keep = _lib._as_ptr(fileContents)
        _lib.check(_lib.lib().rsn_benchmark_file(name.encode(), ptr, n, C.byref(r)))
        del keep
        if r.failed:
            return Result(name, "DNF", float("nan"), float("nan"), float("nan"), False, True)
        return Result(name, f"{r.seconds * 1e3:.2f}ms", float(r.ratio), float(r.actual_entropy), float(r.entropy),
                      bool(r.lossless), False, r.seconds)
    try:
        import numpy as np

        hist = np.bincount(np.frombuffer(fileContents, dtype=np.uint8), minlength=256)
        total = len(fileContents)
        entropy = _entropy(hist.tolist(), total) if total else 0.0
        start = time.perf_counter()
        compressed = (compress_fused if fused else compress)(fileContents, algorithms)
        decompressed = (decompress_fused if fused else decompress)(compressed, algorithms)
        seconds = time.perf_counter() - start
        lossless = decompressed == fileContents
        ratio = len(compressed) / len(fileContents) * 100 if fileContents else float("nan")
        # engine.go:412-423 histograms the DEcompressed content but divides by len(compressed)
        h2 = np.bincount(np.frombuffer(decompressed, dtype=np.uint8), minlength=256)
        actual = _entropy(h2.tolist(), len(compressed)) if compressed else 0.0
        return Result(name, f"{seconds * 1e3:.2f}ms", ratio, actual, entropy, lossless, False, seconds)
    except _lib.RaisinPanic:
        # AsyncBenchmarkFile recovers panics into Failed rows (engine.go:315-328)
        return Result(name, "DNF", float("nan"), float("nan"), float("nan"), False, True)
