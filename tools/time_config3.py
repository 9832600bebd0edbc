#!/usr/bin/env python
"""BASELINE configs[2] timed properly: layered lzss,huffman on a large synthetic mixed corpus, one
stream, (a) device-resident through rsn_dev_* calls, CUDA events, (b) pinned host buffers through
rsn_compress_layers / rsn_decompress_layers, wall clock; three repetitions after a full-size warm-up
(the first call of a size grows the arena and the pinned pool).  Parity at this size is
tests/tools/validate_large.py's job.  RSN_TRACE=1 prints the stages.
usage: python tools/time_config3.py [MiB=1024]"""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import raisin_b200 as rsn  # noqa: E402
from raisin_b200 import synth  # noqa: E402

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = mib << 20
data = synth.mixed(n, 3)
lib = rsn._lib.lib()
rsn._lib.check(lib.rsn_init(0))
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
sp = C.c_void_p(stream.cuda_stream)
h_in = lib.rsn_host_alloc(n)
C.memmove(h_in, data, n)
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
rsn._lib.check(lib.rsn_dev_upload(h_in, n, d_in.data_ptr(), sp))


def dev_once():
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    o1, n1, o2, n2, o3, n3, o4, n4 = (C.c_void_p(), C.c_size_t(), C.c_void_p(), C.c_size_t(), C.c_void_p(), C.c_size_t(),
                                      C.c_void_p(), C.c_size_t())
    ev[0].record()
    rsn._lib.check(lib.rsn_dev_lzss_compress(d_in.data_ptr(), n, 4096, 0, C.byref(o1), C.byref(n1), sp))
    ev[1].record()
    rsn._lib.check(lib.rsn_dev_huff_compress(o1, n1.value, C.byref(o2), C.byref(n2), sp))
    ev[2].record()
    rsn._lib.check(lib.rsn_dev_huff_decompress(o2, n2.value, 0, C.byref(o3), C.byref(n3), sp))
    ev[3].record()
    rsn._lib.check(lib.rsn_dev_lzss_decompress(o3, n3.value, C.byref(o4), C.byref(n4), sp))
    ev[4].record()
    torch.cuda.synchronize()
    for o in (o1, o2, o3, o4):
        lib.rsn_dev_free(o, sp)
    t = [ev[i].elapsed_time(ev[i + 1]) for i in range(4)]
    return t, n1.value, n2.value, n4.value


def host_once():
    o, on = C.POINTER(C.c_uint8)(), C.c_size_t()
    t0 = time.perf_counter()
    rsn._lib.check(lib.rsn_compress_layers(b"lzss,huffman", C.c_void_p(h_in), n, C.byref(o), C.byref(on)))
    t1 = time.perf_counter()
    b, bn = C.POINTER(C.c_uint8)(), C.c_size_t()
    rsn._lib.check(lib.rsn_decompress_layers(b"lzss,huffman", C.cast(o, C.c_void_p), on.value, C.byref(b), C.byref(bn)))
    t2 = time.perf_counter()
    lib.rsn_free(o)
    lib.rsn_free(b)
    return (t1 - t0) * 1e3, (t2 - t1) * 1e3, on.value, bn.value


dev_once()
dev = [dev_once() for _ in range(3)]
host_once()
host = [host_once() for _ in range(3)]
best = min(dev, key=lambda r: sum(r[0]))
hb = min(host, key=lambda r: r[0] + r[1])
print(json.dumps({
    "workload": f"layered lzss,huffman, {mib} MiB synthetic mixed corpus, one stream (BASELINE configs[2])",
    "input_bytes": n, "lzss_bytes": best[1], "compressed_bytes": best[2], "decoded_bytes": best[3],
    "device_resident_ms": {"lzss_compress": best[0][0], "huffman_compress": best[0][1], "huffman_decompress": best[0][2],
                           "lzss_decompress": best[0][3], "total": sum(best[0])},
    "device_resident_GBps": n / (sum(best[0]) / 1e3) / 1e9,
    "host_buffers_ms": {"compress_layers": hb[0], "decompress_layers": hb[1], "total": hb[0] + hb[1]},
    "host_buffers_GBps": n / ((hb[0] + hb[1]) / 1e3) / 1e9,
    "all_device_runs_ms": [[round(x, 2) for x in r[0]] for r in dev],
    "timing": "device: CUDA events on the calls' stream; host: wall clock around the C-ABI calls, pinned buffers; best of 3",
}))
