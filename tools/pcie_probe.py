#!/usr/bin/env python
"""Host<->device copy rates on this box for the shapes the batch path uses: one large copy, many 256 KiB
copies on one stream, the same from several threads/streams, and both directions at once.
usage: python tools/pcie_probe.py"""
import threading
import time

import torch

N = 2048
SZ = 256 << 10
big_h = torch.empty(N * SZ, dtype=torch.uint8).pin_memory()
big_d = torch.empty(N * SZ, dtype=torch.uint8, device="cuda")
small_h = [torch.empty(SZ, dtype=torch.uint8).pin_memory() for _ in range(N)]
small_d = torch.empty(N * SZ, dtype=torch.uint8, device="cuda")
out_h = torch.empty(N * SZ, dtype=torch.uint8).pin_memory()


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


gb = N * SZ / 1e9
t = timed(lambda: big_d.copy_(big_h, non_blocking=True))
print(f"H2D one {N * SZ >> 20} MiB copy: {gb / t:.1f} GB/s")
t = timed(lambda: out_h.copy_(big_d, non_blocking=True))
print(f"D2H one {N * SZ >> 20} MiB copy: {gb / t:.1f} GB/s")


def many(lo, hi, stream):
    with torch.cuda.stream(stream):
        for i in range(lo, hi):
            small_d[i * SZ:(i + 1) * SZ].copy_(small_h[i], non_blocking=True)


s0 = torch.cuda.Stream()
t = timed(lambda: many(0, N, s0))
print(f"H2D {N} x 256 KiB, one thread/stream: {gb / t:.1f} GB/s ({t / N * 1e6:.1f} us per copy)")
for k in (2, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(k)]

    def run():
        th = [threading.Thread(target=many, args=(j * N // k, (j + 1) * N // k, streams[j])) for j in range(k)]
        for x in th:
            x.start()
        for x in th:
            x.join()

    t = timed(run)
    print(f"H2D {N} x 256 KiB, {k} threads/streams: {gb / t:.1f} GB/s")


def both():
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    with torch.cuda.stream(sa):
        big_d.copy_(big_h, non_blocking=True)
    with torch.cuda.stream(sb):
        out_h.copy_(small_d, non_blocking=True)


t = timed(both)
print(f"H2D + D2H of {N * SZ >> 20} MiB each, concurrently: {2 * gb / t:.1f} GB/s total")
