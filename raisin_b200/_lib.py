"""ctypes loader for libraisin_b200.so (the C ABI in include/raisin_b200.h).

There is no fallback: if the shared library is missing this raises, and every codec call
fails with RSN_ERR_NO_DEVICE / RSN_ERR_CUDA when no CUDA device is usable.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("RSN_LIB_PATH") or os.path.join(_HERE, "libraisin_b200.so")  # RSN_LIB_PATH: A/B runs of two builds

# every symbol include/raisin_b200.h declares
EXPORTS = [
    "rsn_init", "rsn_shutdown", "rsn_strerror", "rsn_last_cuda_error", "rsn_free", "rsn_free_many", "rsn_dev_free_many", "rsn_host_alloc",
    "rsn_host_free", "rsn_lzss_compress", "rsn_lzss_compress_sharded", "rsn_sharded_peer_bytes", "rsn_lzss_decompress", "rsn_huff_compress", "rsn_huff_decompress",
    "rsn_compress_layers", "rsn_decompress_layers", "rsn_benchmark_file", "rsn_batch_layers", "rsn_batch_plan", "rsn_dev_lzss_compress", "rsn_dev_lzss_decompress",
    "rsn_dev_huff_compress", "rsn_dev_huff_decompress", "rsn_dev_free", "rsn_dev_download", "rsn_dev_upload",
    "rsn_dev_lzss_match", "rsn_dev_lzss_emit", "rsn_dev_lzss_escape",
    "rsn_kernel_launches", "rsn_reset_kernel_launches", "rsn_kernel_timing", "rsn_kernel_timing_report", "rsn_version",
]

RSN_LZSS_ASYNC = 0
RSN_LZSS_ITER = 1


class RaisinPanic(RuntimeError):
    """The reference signals these conditions by panicking; the shim does the same."""

    def __init__(self, rc: int, msg: str):
        super().__init__(f"raisin_b200: {msg} (rc={rc})")
        self.rc = rc


class BenchResult(C.Structure):
    _fields_ = [("seconds", C.c_double), ("entropy", C.c_double), ("actual_entropy", C.c_float), ("ratio", C.c_float),
                ("lossless", C.c_int), ("failed", C.c_int), ("error", C.c_int), ("compressed_n", C.c_size_t),
                ("decompressed_n", C.c_size_t)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C raisin_b200/csrc). There is no CPU fallback."
        )
    L = C.CDLL(SO_PATH)
    u8pp = C.POINTER(C.POINTER(C.c_uint8))
    szp = C.POINTER(C.c_size_t)
    L.rsn_init.argtypes = [C.c_int]
    L.rsn_init.restype = C.c_int
    L.rsn_shutdown.argtypes = []
    L.rsn_shutdown.restype = None
    L.rsn_strerror.argtypes = [C.c_int]
    L.rsn_strerror.restype = C.c_char_p
    L.rsn_last_cuda_error.argtypes = []
    L.rsn_last_cuda_error.restype = C.c_char_p
    L.rsn_free.argtypes = [C.c_void_p]
    L.rsn_free.restype = None
    L.rsn_free_many.argtypes = [C.c_void_p, C.c_size_t]
    L.rsn_free_many.restype = None
    L.rsn_dev_free_many.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
    L.rsn_dev_free_many.restype = None
    L.rsn_host_alloc.argtypes = [C.c_size_t]
    L.rsn_host_alloc.restype = C.c_void_p
    L.rsn_host_free.argtypes = [C.c_void_p]
    L.rsn_host_free.restype = None
    L.rsn_lzss_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_int64, C.c_int, u8pp, szp]
    L.rsn_lzss_compress_sharded.argtypes = [C.c_void_p, C.c_size_t, C.c_int64, C.c_int, C.c_int, u8pp, szp]
    L.rsn_lzss_compress_sharded.restype = C.c_int
    L.rsn_sharded_peer_bytes.argtypes = []
    L.rsn_sharded_peer_bytes.restype = C.c_uint64
    L.rsn_lzss_decompress.argtypes = [C.c_void_p, C.c_size_t, u8pp, szp]
    L.rsn_huff_compress.argtypes = [C.c_void_p, C.c_size_t, u8pp, szp]
    L.rsn_huff_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_int, u8pp, szp]
    L.rsn_compress_layers.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t, u8pp, szp]
    L.rsn_decompress_layers.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t, u8pp, szp]
    L.rsn_benchmark_file.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t, C.POINTER(BenchResult)]
    L.rsn_benchmark_file.restype = C.c_int
    L.rsn_batch_layers.argtypes = [C.c_char_p, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_int, C.c_int]
    L.rsn_batch_layers.restype = C.c_int
    L.rsn_batch_plan.argtypes = [C.c_size_t, C.c_void_p, C.c_void_p, szp]
    L.rsn_batch_plan.restype = C.c_int
    vpp = C.POINTER(C.c_void_p)
    L.rsn_dev_lzss_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_int64, C.c_int, vpp, szp, C.c_void_p]
    L.rsn_dev_lzss_decompress.argtypes = [C.c_void_p, C.c_size_t, vpp, szp, C.c_void_p]
    L.rsn_dev_huff_compress.argtypes = [C.c_void_p, C.c_size_t, vpp, szp, C.c_void_p]
    L.rsn_dev_huff_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_int, vpp, szp, C.c_void_p]
    L.rsn_dev_free.argtypes = [C.c_void_p, C.c_void_p]
    L.rsn_dev_free.restype = None
    L.rsn_dev_download.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    L.rsn_dev_upload.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    L.rsn_dev_lzss_emit.argtypes = [C.c_void_p, C.c_size_t, C.c_int64, C.c_int, C.c_void_p, vpp, szp, C.c_void_p]
    L.rsn_dev_lzss_escape.argtypes = [C.c_void_p, C.c_size_t, vpp, szp, C.c_void_p]
    L.rsn_dev_lzss_match.argtypes = [C.c_void_p, C.c_size_t, C.c_int64, C.c_void_p, C.c_void_p]
    for name in ("rsn_lzss_compress", "rsn_lzss_decompress", "rsn_huff_compress", "rsn_huff_decompress",
                 "rsn_compress_layers", "rsn_decompress_layers", "rsn_dev_lzss_compress", "rsn_dev_lzss_decompress",
                 "rsn_dev_huff_compress", "rsn_dev_huff_decompress", "rsn_dev_lzss_match", "rsn_dev_lzss_emit", "rsn_dev_lzss_escape", "rsn_dev_download",
                 "rsn_dev_upload"):
        getattr(L, name).restype = C.c_int
    L.rsn_kernel_launches.argtypes = []
    L.rsn_kernel_launches.restype = C.c_uint64
    L.rsn_reset_kernel_launches.argtypes = []
    L.rsn_reset_kernel_launches.restype = None
    L.rsn_kernel_timing.argtypes = [C.c_int]
    L.rsn_kernel_timing.restype = None
    L.rsn_kernel_timing_report.argtypes = [C.c_char_p, C.c_size_t]
    L.rsn_kernel_timing_report.restype = C.c_size_t
    L.rsn_version.argtypes = []
    L.rsn_version.restype = C.c_char_p
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        L = lib()
        msg = L.rsn_strerror(rc).decode()
        if rc in (-1, -2):
            detail = L.rsn_last_cuda_error().decode()
            if detail:
                msg += ": " + detail
        raise RaisinPanic(rc, msg)


def _as_ptr(data):
    """(pointer, length, keepalive) for bytes / bytearray / memoryview / numpy uint8 arrays."""
    if isinstance(data, (bytes, bytearray)):
        n = len(data)
        if isinstance(data, bytes):
            return C.cast(C.c_char_p(data), C.c_void_p), n, data
        buf = (C.c_uint8 * n).from_buffer(data) if n else None
        return C.cast(buf, C.c_void_p) if n else C.c_void_p(0), n, (data, buf)
    try:
        import numpy as np

        if isinstance(data, np.ndarray):
            a = np.ascontiguousarray(data, dtype=np.uint8)
            return C.c_void_p(a.ctypes.data), a.size, a
    except ImportError:
        pass
    mv = memoryview(data).cast("B")
    b = mv.tobytes()
    return C.cast(C.c_char_p(b), C.c_void_p), len(b), b


def bytes_at(ptr, n: int) -> bytes:
    """Copy n bytes from a C pointer (ctypes.string_at takes a C int: results of 2 GiB and more
    need the buffer route)."""
    if n < (1 << 31) - 1:
        return C.string_at(ptr, n)
    addr = C.cast(ptr, C.c_void_p).value
    return bytes((C.c_ubyte * n).from_address(addr))


def call_host(fn, data, *extra) -> bytes:
    """Run a host-buffer entry point `fn(in, n, *extra, &out, &out_n)` and copy the result."""
    L = lib()
    ptr, n, keep = _as_ptr(data)
    out = C.POINTER(C.c_uint8)()
    out_n = C.c_size_t(0)
    rc = fn(ptr, n, *extra, C.byref(out), C.byref(out_n))
    del keep
    check(rc)
    try:
        return bytes_at(out, out_n.value)
    finally:
        L.rsn_free(out)
