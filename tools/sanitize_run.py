#!/usr/bin/env python
"""Small end-to-end run of every kernel for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import raisin_b200 as rsn  # noqa: E402
from raisin_b200 import synth  # noqa: E402

datas = [synth.mixed(70000, 3, segment=12000), b"a" * 9001, synth.text(20001, 4), b"<\\\xff" * 700 + b"xyz", b"q"]
for d in datas:
    for w in (4096, 100, 9000):
        a = rsn.lz.CompressAsync(d, False, w)
        assert rsn.lz.Decompress(a) == d
        b = rsn.lz.Compress(d, False, w)
        try:
            rsn.lz.Decompress(b)
        except rsn.RaisinPanic:
            pass
    h = rsn.huffman.Compress(d)
    try:
        rsn.huffman.Decompress(h)
    except rsn.RaisinPanic:
        pass
    x = rsn.engine.compress_fused(d, ["lzss", "huffman"])
    try:
        rsn.engine.decompress_fused(x, ["lzss", "huffman"])
    except rsn.RaisinPanic:
        pass
# the batched small-file path (kb_* kernels, device tree builder), both directions, incl. junk streams
files = datas + [synth.batch_file(j, 30000) for j in range(6)] + [b"", "".join(chr(0x100 + i) for i in range(900)).encode()]
for algos in (["lzss"], ["huffman"], ["lzss", "huffman"]):
    got = rsn.engine.batch(files, algos, True, workers=2)
    streams = [g if g is not None else b"junk<9,4>" for g in got] + [b"abc<9,4>def", b"no separator", b"3|a1|b\\\n\x00\xff"]
    rsn.engine.batch(streams, algos, False, workers=2)
print("sanitize run ok")
