#!/usr/bin/env python
"""Per-kernel device time of one batch pass (compress + decompress, one worker so kernels run alone)
for each kind of config-4 file on its own: text (j % 3 == 0), logs (1), random bytes (2).
usage: python tools/batch_kinds.py [files per kind = 256]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

per = int(sys.argv[1]) if len(sys.argv) > 1 else 256
sets = {k: bench.make_files([3 * i + k for i in range(per)]) for k in range(3)}
import torch  # noqa: E402

import raisin_b200 as rsn  # noqa: E402

lib = rsn._lib.lib()
rsn._lib.check(lib.rsn_init(0))
for k, name in enumerate(("text", "logs", "random")):
    files = sets[k]
    n = len(files)
    ns = (C.c_size_t * n)(*[len(f) for f in files])
    d_blob = torch.frombuffer(bytearray(b"".join(files)), dtype=torch.uint8).cuda()
    ins = (C.c_void_p * n)(*[d_blob.data_ptr() + 262144 * i for i in range(n)])

    def one():
        outs, out_ns, rcs = (C.c_void_p * n)(), (C.c_size_t * n)(), (C.c_int * n)()
        rsn._lib.check(lib.rsn_batch_layers(b"lzss,huffman", 1, n, ins, ns, outs, out_ns, rcs, 1, 1))
        b_outs, b_ns = (C.c_void_p * n)(), (C.c_size_t * n)()
        rsn._lib.check(lib.rsn_batch_layers(b"lzss,huffman", 0, n, outs, out_ns, b_outs, b_ns, rcs, 1, 1))
        c = sum(out_ns)
        lib.rsn_dev_free_many(outs, n, None)
        lib.rsn_dev_free_many(b_outs, n, None)
        return c

    for _ in range(2):
        one()
    lib.rsn_kernel_timing(1)
    c = one()
    rows = bench.kernel_report(lib)
    lib.rsn_kernel_timing(0)
    tot = sum(r[2] for r in rows)
    print(f"== {name}: {n} files, {sum(ns) >> 20} MiB -> {c >> 20} MiB; kernel time {tot:.2f} ms")
    for r in rows[:14]:
        print(f"   {r[0]:24s} {r[1]:4d} launches {r[2]:8.3f} ms  ({sum(ns) / 1e6 / r[2]:8.1f} GB/s of input)")
