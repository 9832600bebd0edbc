#!/usr/bin/env python
"""Small driver for ncu captures: runs one stage of the path a few times on synthetic data.
usage: python tools/prof_run.py {match|lzc|lzd|hc|hd} [MiB] [kind] [iters]"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import raisin_b200 as rsn  # noqa: E402
from raisin_b200 import synth  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "match"
mib = float(sys.argv[2]) if len(sys.argv) > 2 else 8
kind = sys.argv[3] if len(sys.argv) > 3 else "text"
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
n = int(mib * (1 << 20))
lib = rsn._lib.lib()
rsn._lib.check(lib.rsn_init(0))
data = synth.generate(kind, n, 2)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
sp = C.c_void_p(stream.cuda_stream)
d_in = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
torch.cuda.synchronize()


def run(fn_in, n_in):
    out, on = C.c_void_p(), C.c_size_t()
    if what in ("lzc",):
        rsn._lib.check(lib.rsn_dev_lzss_compress(fn_in, n_in, 4096, 0, C.byref(out), C.byref(on), sp))
    elif what == "lzd":
        rsn._lib.check(lib.rsn_dev_lzss_decompress(fn_in, n_in, C.byref(out), C.byref(on), sp))
    elif what == "hc":
        rsn._lib.check(lib.rsn_dev_huff_compress(fn_in, n_in, C.byref(out), C.byref(on), sp))
    elif what == "hd":
        rsn._lib.check(lib.rsn_dev_huff_decompress(fn_in, n_in, 0, C.byref(out), C.byref(on), sp))
    return out, on.value


src, src_n = d_in.data_ptr(), n
keep = None
if what == "lzd":
    o, c = C.c_void_p(), C.c_size_t()
    rsn._lib.check(lib.rsn_dev_lzss_compress(src, n, 4096, 0, C.byref(o), C.byref(c), sp))
    src, src_n = o, c.value
elif what == "hd":
    o, c = C.c_void_p(), C.c_size_t()
    rsn._lib.check(lib.rsn_dev_huff_compress(src, n, C.byref(o), C.byref(c), sp))
    src, src_n = o, c.value
packed = torch.empty(n, dtype=torch.int32, device="cuda")
for it in range(iters):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if what == "match":
        rsn._lib.check(lib.rsn_dev_lzss_match(src, n, 4096, packed.data_ptr(), sp))
    else:
        o, on = run(src, src_n)
        lib.rsn_dev_free(o, sp)
    torch.cuda.synchronize()
    print(f"{what} {kind} {mib} MiB iter {it}: {(time.perf_counter() - t0) * 1e3:.3f} ms", flush=True)
