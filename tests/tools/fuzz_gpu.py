#!/usr/bin/env python
"""Randomised parity run: GPU path vs oracle on structured random inputs (all ops, both LZSS
variants, several windows).  usage: python tests/tools/fuzz_gpu.py [seconds=120] [seed=0]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import raisin_b200 as rsn  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from raisin_b200 import synth  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rng = np.random.default_rng(seed)
ALPHAS = [b"ab", b"abc<\\\xff", b"0123456789,<>", bytes(range(256)), b"the quick brown fox ", "aé☃😀|\\\n".encode(),
          bytes([0x5C]), b"<", b"\xff\\"]


def gen():
    kind = int(rng.integers(0, 8))
    n = int(rng.integers(0, 60000)) if rng.random() < 0.7 else int(rng.integers(0, 400))
    if kind == 0:
        return synth.text(n, int(rng.integers(1, 1 << 30)))
    if kind == 1:
        return synth.logs(n, int(rng.integers(1, 1 << 30)))
    if kind == 2:
        return synth.random_bytes(n, int(rng.integers(1, 1 << 30)))
    if kind == 3:
        a = ALPHAS[int(rng.integers(0, len(ALPHAS)))]
        return bytes(a[i] for i in rng.integers(0, len(a), size=min(n, 20000)))
    if kind == 4:  # periodic with mutations
        p = int(rng.integers(1, 5000))
        base = synth.random_bytes(p, int(rng.integers(1, 1 << 30)))
        buf = bytearray((base * (n // p + 1))[:n])
        for _ in range(int(rng.integers(0, 20))):
            if n:
                buf[int(rng.integers(0, n))] = int(rng.integers(0, 256))
        return bytes(buf)
    if kind == 5:  # long runs
        out = bytearray()
        while len(out) < n:
            out += bytes([int(rng.integers(0x61, 0x64))]) * int(rng.integers(1, 9000))
        return bytes(out[:n])
    if kind == 6:  # token-like junk
        a = b"<>,0123456789+-ab"
        return bytes(a[i] for i in rng.integers(0, len(a), size=min(n, 5000)))
    return synth.mixed(n, int(rng.integers(1, 1 << 30)), segment=max(1, n // 5))


t_end = time.time() + budget
cases = 0
while time.time() < t_end:
    d = gen()
    w = int(rng.choice([4096, 4096, 4096, 1, 2, 7, 100, 1000, 4095, 5000, 20000, 0]))
    if w == 0 and len(d) > 30000:
        w = 20000  # unbounded windows above 32768 bytes are a documented limit (RSN_ERR_UNSUPPORTED)
    tag = (cases, len(d), w)
    a = rsn.lz.CompressAsync(d, False, w)
    assert a == po.lzss_compress_async(d, w, threads=4), ("async", tag)
    assert rsn.lz.Decompress(a) == d, ("roundtrip", tag)
    b = rsn.lz.Compress(d, False, w)
    assert b == po.lzss_compress_iter(d, w), ("iter", tag)
    for stream in (b, d):  # variant-B streams and the raw input taken as a (mostly malformed) stream
        try:
            want = po.lzss_decompress(stream)
        except po.OracleError:
            try:
                rsn.lz.Decompress(stream)
                raise AssertionError(("decoder accepted a stream the reference rejects", tag))
            except rsn.RaisinPanic as e:
                assert e.rc == -15, (e.rc, tag)
        else:
            assert rsn.lz.Decompress(stream) == want, ("decode arbitrary", tag)
    if d:
        h = rsn.huffman.Compress(d)
        assert h == po.huff_compress(d), ("huff", tag)
        try:
            want = po.huff_decompress(h)
        except po.OracleError as oe:
            try:
                rsn.huffman.Decompress(h)
                raise AssertionError(("huff decoder accepted", tag))
            except rsn.RaisinPanic:
                pass
        else:
            assert rsn.huffman.Decompress(h) == want, ("huff decode", tag)
        lay = rsn.engine.compress_fused(d, ["lzss", "huffman"])
        assert lay == po.huff_compress(po.lzss_compress_async(d, 4096, threads=4)), ("layered", tag)
    cases += 1
print(f"fuzz ok: {cases} cases, seed {seed}")
