#!/usr/bin/env python
"""Timing probe for rsn_batch_layers: files of one kind (or the config-4 mix), several worker counts.
usage: python tools/batch_probe.py [kind|mix] [files] [size] [algos]"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raisin_b200 as rsn  # noqa: E402
from raisin_b200 import synth  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "mix"
nfiles = int(sys.argv[2]) if len(sys.argv) > 2 else 512
size = int(sys.argv[3]) if len(sys.argv) > 3 else 262144
algos = (sys.argv[4] if len(sys.argv) > 4 else "lzss,huffman").encode()
lib = rsn._lib.lib()
rsn._lib.check(lib.rsn_init(0))
files = [synth.batch_file(j, size) if kind == "mix" else synth.generate(kind, size, 1000 + j) for j in range(nfiles)]
n = len(files)
total = sum(len(f) for f in files)
keep = [rsn._lib._as_ptr(f) for f in files]
ins = (C.c_void_p * n)(*[k[0] for k in keep])
if os.environ.get("PROBE_PINNED"):  # inputs in pinned host memory instead of pageable Python bytes
    lib.rsn_host_alloc.restype = C.c_void_p
    pinned = []
    for f in files:
        p = lib.rsn_host_alloc(C.c_size_t(len(f)))
        C.memmove(p, f, len(f))
        pinned.append(p)
    ins = (C.c_void_p * n)(*pinned)
ns = (C.c_size_t * n)(*[k[1] for k in keep])
for workers in (1, 2, 4, 6, 8, 12):
    best_c = best_d = 1e9
    for it in range(5):
        outs, out_ns, rcs = (C.c_void_p * n)(), (C.c_size_t * n)(), (C.c_int * n)()
        t0 = time.perf_counter()
        rsn._lib.check(lib.rsn_batch_layers(algos, 1, n, ins, ns, outs, out_ns, rcs, workers, 0))
        t1 = time.perf_counter()
        b_outs, b_ns = (C.c_void_p * n)(), (C.c_size_t * n)()
        rc = lib.rsn_batch_layers(algos, 0, n, outs, out_ns, b_outs, b_ns, rcs, workers, 0)
        t2 = time.perf_counter()
        for i in range(n):
            lib.rsn_free(outs[i])
            lib.rsn_free(b_outs[i])
        if it:
            best_c, best_d = min(best_c, t1 - t0), min(best_d, t2 - t1)
    print(f"{kind} {n} x {size} B {algos.decode()} workers {workers}: compress {best_c * 1e3:.1f} ms "
          f"({total / best_c / 1e9:.2f} GB/s), decompress {best_d * 1e3:.1f} ms ({total / best_d / 1e9:.2f} GB/s)", flush=True)
