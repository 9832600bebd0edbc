import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import raisin_b200 as rsn
from raisin_b200 import synth
lib = rsn._lib.lib(); rsn._lib.check(lib.rsn_init(0))
n = 64 << 20
data = synth.text(n, 2)
h_in = lib.rsn_host_alloc(n); C.memmove(h_in, data, n)
def t(fn, reps=6):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts), r
last = []
def comp():
    if last:
        lib.rsn_free(last.pop())
    o = C.POINTER(C.c_uint8)(); on = C.c_size_t()
    rsn._lib.check(lib.rsn_lzss_compress(h_in, n, 4096, 0, C.byref(o), C.byref(on)))
    last.append(o)
    return o, on.value
ms, (o, on) = t(comp)
print(f"host compress   {ms:.2f} ms")
def dec():
    b = C.POINTER(C.c_uint8)(); bn = C.c_size_t()
    rsn._lib.check(lib.rsn_lzss_decompress(o, on, C.byref(b), C.byref(bn)))
    lib.rsn_free(b); return None
ms, _ = t(dec)
print(f"host decompress {ms:.2f} ms")
# raw copy speeds
d = torch.empty(n, dtype=torch.uint8, device="cuda")
rt = torch.cuda.current_stream()
import numpy as np
hp = torch.empty(n, dtype=torch.uint8).pin_memory()
ms, _ = t(lambda: (d.copy_(hp, non_blocking=True), torch.cuda.synchronize()))
print(f"H2D 64 MiB pinned {ms:.2f} ms")
ms, _ = t(lambda: (hp.copy_(d, non_blocking=True), torch.cuda.synchronize()))
print(f"D2H 64 MiB pinned {ms:.2f} ms")
