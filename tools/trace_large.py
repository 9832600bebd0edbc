import sys, time, os, ctypes as C
sys.path.insert(0, '/root/repo')
import torch
import raisin_b200 as rsn
from raisin_b200 import synth
lib = rsn._lib.lib(); rsn._lib.check(lib.rsn_init(0))
n = 256 << 20
data = synth.mixed(n, 3)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); sp = C.c_void_p(stream.cuda_stream)
d_in = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
def t(label, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); print(f"{label}: {(time.perf_counter()-t0)*1e3:.2f} ms", flush=True); return r
def lzc(src, sn):
    o, on = C.c_void_p(), C.c_size_t(); rsn._lib.check(lib.rsn_dev_lzss_compress(src, sn, 4096, 0, C.byref(o), C.byref(on), sp)); return o, on.value
def hc(src, sn):
    o, on = C.c_void_p(), C.c_size_t(); rsn._lib.check(lib.rsn_dev_huff_compress(src, sn, C.byref(o), C.byref(on), sp)); return o, on.value
def hd(src, sn):
    o, on = C.c_void_p(), C.c_size_t(); rsn._lib.check(lib.rsn_dev_huff_decompress(src, sn, 0, C.byref(o), C.byref(on), sp)); return o, on.value
def lzd(src, sn):
    o, on = C.c_void_p(), C.c_size_t(); rsn._lib.check(lib.rsn_dev_lzss_decompress(src, sn, C.byref(o), C.byref(on), sp)); return o, on.value
for it in range(3):
    a, an = t("lzc", lambda: lzc(d_in.data_ptr(), n))
    b, bn = t("hc ", lambda: hc(a, an))
    c, cn = t("hd ", lambda: hd(b, bn))
    d, dn = t("lzd", lambda: lzd(c, cn))
    for p in (a, b, c, d): lib.rsn_dev_free(p, sp)
    print(an, bn, cn, dn)
