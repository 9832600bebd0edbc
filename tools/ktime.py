#!/usr/bin/env python
"""Per-kernel device times (library launch timer) of one call: python tools/ktime.py {lzc|lzd|hc|hd} [MiB] [kind]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import raisin_b200 as rsn  # noqa: E402
from raisin_b200 import synth  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "lzc"
mib = float(sys.argv[2]) if len(sys.argv) > 2 else 64
kind = sys.argv[3] if len(sys.argv) > 3 else "text"
n = int(mib * (1 << 20))
lib = rsn._lib.lib()
rsn._lib.check(lib.rsn_init(0))
data = synth.generate(kind, n, 2)
d_in = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
sp = None


def call(fn, src, sn, *extra):
    o, on = C.c_void_p(), C.c_size_t()
    rsn._lib.check(fn(src, sn, *extra, C.byref(o), C.byref(on), sp))
    return o, on.value


def run():
    if what == "lzc":
        o, _ = call(lib.rsn_dev_lzss_compress, d_in.data_ptr(), n, 4096, 0)
        lib.rsn_dev_free(o, sp)
    elif what == "lzd":
        o, _ = call(lib.rsn_dev_lzss_decompress, src, src_n)
        lib.rsn_dev_free(o, sp)
    elif what == "hc":
        o, _ = call(lib.rsn_dev_huff_compress, d_in.data_ptr(), n)
        lib.rsn_dev_free(o, sp)
    elif what == "hd":
        o, _ = call(lib.rsn_dev_huff_decompress, src, src_n, 0)
        lib.rsn_dev_free(o, sp)


if what == "lzd":
    src, src_n = call(lib.rsn_dev_lzss_compress, d_in.data_ptr(), n, 4096, 0)
elif what == "hd":
    src, src_n = call(lib.rsn_dev_huff_compress, d_in.data_ptr(), n)
for _ in range(2):
    run()
lib.rsn_kernel_timing(1)
run()
k = lib.rsn_kernel_timing_report(None, 0)
buf = C.create_string_buffer(k + 1)
lib.rsn_kernel_timing_report(buf, k + 1)
lib.rsn_kernel_timing(0)
print(f"{what} {kind} {mib} MiB")
print(buf.value.decode())
