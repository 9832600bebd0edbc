#!/usr/bin/env python
"""Kernel timeline of one batch compress + decompress (library launch timer, RSN_KTIME_TIMELINE):
how much of the wall time has 0, 1, 2+ kernels in flight, and which kernels run alone.
usage: python tools/batch_timeline.py [files] [workers] [host]   (host: pinned host buffers in and out instead of device buffers)"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["RSN_KTIME_TIMELINE"] = "/tmp/rsn_timeline.txt"
import bench  # noqa: E402

nfiles = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
workers = int(sys.argv[2]) if len(sys.argv) > 2 else 8
host_mode = len(sys.argv) > 3 and sys.argv[3] == "host"
files = bench.make_files(list(range(nfiles)))
import torch  # noqa: E402

import raisin_b200 as rsn  # noqa: E402

lib = rsn._lib.lib()
rsn._lib.check(lib.rsn_init(0))
n = len(files)
ns = (C.c_size_t * n)(*[len(f) for f in files])
d_blob = torch.frombuffer(bytearray(b"".join(files)), dtype=torch.uint8).cuda()
ins = (C.c_void_p * n)(*[d_blob.data_ptr() + 262144 * i for i in range(n)])
if host_mode:
    h_ptrs = []
    for f in files:
        p = lib.rsn_host_alloc(len(f))
        C.memmove(p, f, len(f))
        h_ptrs.append(p)
    ins = (C.c_void_p * n)(*h_ptrs)
dev = 0 if host_mode else 1


def one(timing):
    outs, out_ns, rcs = (C.c_void_p * n)(), (C.c_size_t * n)(), (C.c_int * n)()
    if timing:
        lib.rsn_kernel_timing(1)
    t0 = time.perf_counter()
    rsn._lib.check(lib.rsn_batch_layers(b"lzss,huffman", 1, n, ins, ns, outs, out_ns, rcs, workers, dev))
    t1 = time.perf_counter()
    b_outs, b_ns = (C.c_void_p * n)(), (C.c_size_t * n)()
    rsn._lib.check(lib.rsn_batch_layers(b"lzss,huffman", 0, n, outs, out_ns, b_outs, b_ns, rcs, workers, dev))
    t2 = time.perf_counter()
    if timing:
        k = lib.rsn_kernel_timing_report(None, 0)
        lib.rsn_kernel_timing(0)
    if host_mode:
        lib.rsn_free_many(outs, n)
        lib.rsn_free_many(b_outs, n)
    else:
        lib.rsn_dev_free_many(outs, n, None)
        lib.rsn_dev_free_many(b_outs, n, None)
    return (t1 - t0) * 1e3, (t2 - t1) * 1e3


for _ in range(3):
    one(False)
if os.environ.get("TIMELINE_REPEAT"):  # spread of the untimed passes
    for k in range(int(os.environ["TIMELINE_REPEAT"])):
        print("pass %d: compress %.1f ms, decompress %.1f ms" % ((k,) + one(False)), flush=True)
print("untimed: compress %.1f ms, decompress %.1f ms" % one(False))
c, d = one(True)
print("with launch timer: compress %.1f ms, decompress %.1f ms" % (c, d))
rows = []
for ln in open("/tmp/rsn_timeline.txt"):
    name, st, a, b = ln.rsplit(" ", 3)
    rows.append((name, st, float(a), float(b)))
t_end = max(r[3] for r in rows)
ev = []
for r in rows:
    ev.append((r[2], 1, r[0]))
    ev.append((r[3], -1, r[0]))
ev.sort()
depth = 0
last = 0.0
hist = {}
alone = {}
active = {}
for t, d, name in ev:
    hist[min(depth, 4)] = hist.get(min(depth, 4), 0.0) + (t - last)
    if depth == 1:
        k = next(iter(active))
        alone[k] = alone.get(k, 0.0) + (t - last)
    last = t
    depth += d
    if d > 0:
        active[name] = active.get(name, 0) + 1
    else:
        active[name] -= 1
        if active[name] == 0:
            del active[name]
print(f"span {t_end:.1f} ms; time with k kernels in flight:", {k: round(v, 1) for k, v in sorted(hist.items())})
print("alone:", sorted(((round(v, 1), k) for k, v in alone.items()), reverse=True)[:8])
tot = {}
for r in rows:
    tot[r[0]] = tot.get(r[0], 0.0) + r[3] - r[2]
print("event-time per kernel:", sorted(((round(v, 1), k) for k, v in tot.items()), reverse=True)[:8])
if os.environ.get("TIMELINE_DUMP"):
    streams = sorted({r[1] for r in rows})
    sid = {s: i for i, s in enumerate(streams)}
    for r in sorted(rows, key=lambda r: r[2]):
        if r[3] - r[2] > 0.3 or "tree" in r[0]:
            print(f"{r[2]:8.2f} {r[3]:8.2f}  s{sid[r[1]]}  {r[0]}")
