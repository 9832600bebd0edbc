#!/usr/bin/env python
"""Probe of rsn_batch_layers on the config-4 mix for several worker counts, host or device buffers.
usage: python tools/batch_probe2.py [files] [host|dev] [workers,workers,...]"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

nfiles = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
mode = sys.argv[2] if len(sys.argv) > 2 else "dev"
wl = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "2,4,8").split(",")]
files = bench.make_files(list(range(nfiles)))
import torch  # noqa: E402

import raisin_b200 as rsn  # noqa: E402

lib = rsn._lib.lib()
rsn._lib.check(lib.rsn_init(0))
n = len(files)
total = sum(len(f) for f in files)
ns = (C.c_size_t * n)(*[len(f) for f in files])
if mode == "dev":
    d_blob = torch.frombuffer(bytearray(b"".join(files)), dtype=torch.uint8).cuda()
    ins = (C.c_void_p * n)(*[d_blob.data_ptr() + 262144 * i for i in range(n)])
else:
    ptrs = []
    for f in files:
        p = lib.rsn_host_alloc(len(f))
        C.memmove(p, f, len(f))
        ptrs.append(p)
    ins = (C.c_void_p * n)(*ptrs)
dev = 1 if mode == "dev" else 0
for workers in wl:
    best_c = best_d = 1e9
    for it in range(4):
        outs, out_ns, rcs = (C.c_void_p * n)(), (C.c_size_t * n)(), (C.c_int * n)()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rsn._lib.check(lib.rsn_batch_layers(b"lzss,huffman", 1, n, ins, ns, outs, out_ns, rcs, workers, dev))
        t1 = time.perf_counter()
        b_outs, b_ns = (C.c_void_p * n)(), (C.c_size_t * n)()
        rsn._lib.check(lib.rsn_batch_layers(b"lzss,huffman", 0, n, outs, out_ns, b_outs, b_ns, rcs, workers, dev))
        t2 = time.perf_counter()
        if dev:
            lib.rsn_dev_free_many(outs, n, None)
            lib.rsn_dev_free_many(b_outs, n, None)
        else:
            lib.rsn_free_many(outs, n)
            lib.rsn_free_many(b_outs, n)
        if it:
            best_c, best_d = min(best_c, t1 - t0), min(best_d, t2 - t1)
    print(f"{mode} {n} files group={os.environ.get('RSN_BATCH_GROUP_MIB', '16')}MiB/{os.environ.get('RSN_BATCH_GROUP_FILES', '512')} "
          f"workers {workers}: compress {best_c * 1e3:.1f} ms ({total / best_c / 1e9:.2f} GB/s), decompress "
          f"{best_d * 1e3:.1f} ms ({total / best_d / 1e9:.2f} GB/s), both {total / (best_c + best_d) / 1e9:.2f} GB/s", flush=True)
