#!/usr/bin/env python
"""Randomised check of the batched small-file path: rsn_batch_layers over random groups of files
(valid inputs, escape-heavy junk, corrupted and arbitrary "compressed" streams) must give, file by
file, exactly what the single-stream C-ABI calls give (those are checked against the oracle by
tests/tools/fuzz_gpu.py and the parity tests).  usage: python tools/fuzz_batch.py [seconds=90] [seed=0]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import raisin_b200 as rsn  # noqa: E402
from raisin_b200 import synth  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 90
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rng = np.random.default_rng(seed)
ALPHAS = [b"ab", b"abc<\\\xff", b"0123456789,<>", bytes(range(256)), b"the quick brown fox ", "aé☃😀|\\\n".encode(),
          bytes([0x5C]), b"<", b"\xff\\", b"|0123456789\\n"]


def gen():
    kind = int(rng.integers(0, 8))
    n = int(rng.integers(0, 90000)) if rng.random() < 0.7 else int(rng.integers(0, 400))
    s = int(rng.integers(1, 1 << 30))
    if kind == 0:
        return synth.text(n, s)
    if kind == 1:
        return synth.logs(n, s)
    if kind == 2:
        return synth.random_bytes(n, s)
    if kind == 3:
        a = ALPHAS[int(rng.integers(0, len(ALPHAS)))]
        return bytes(a[i] for i in rng.integers(0, len(a), size=min(n, 20000)))
    if kind == 4:
        p = int(rng.integers(1, 5000))
        base = synth.random_bytes(p, s)
        return (base * (n // p + 1))[:n]
    if kind == 5:
        out = bytearray()
        while len(out) < n:
            out += bytes([int(rng.integers(0x61, 0x64))]) * int(rng.integers(1, 9000))
        return bytes(out[:n])
    if kind == 6:
        a = b"<>,0123456789+-ab"
        return bytes(a[i] for i in rng.integers(0, len(a), size=min(n, 5000)))
    return "".join(chr(int(c)) for c in rng.integers(0x20, 0x3000, size=min(n, 8000))).encode()


def corrupt(b):
    if not b or rng.random() < 0.5:
        return b
    b = bytearray(b)
    for _ in range(int(rng.integers(1, 4))):
        how = int(rng.integers(0, 3))
        if how == 0 and len(b) > 1:
            del b[int(rng.integers(0, len(b))):]
        elif how == 1 and b:
            b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        elif b:
            i = int(rng.integers(0, len(b)))
            b[i:i] = bytes(rng.integers(0, 256, size=int(rng.integers(1, 8)), dtype=np.uint8))
    return bytes(b)


def single(fn, x):
    try:
        return fn(x)
    except rsn.RaisinPanic:
        return None


t_end = time.time() + budget
rounds = files_checked = 0
while time.time() < t_end:
    files = [gen() for _ in range(int(rng.integers(1, 40)))]
    workers = int(rng.integers(1, 5))
    for algos, comp, dec in ((["lzss"], lambda x: rsn.lz.CompressAsync(x, False, 4096), lambda x: rsn.lz.Decompress(x, False)),
                             (["huffman"], rsn.huffman.Compress, rsn.huffman.Decompress),
                             (["lzss", "huffman"], lambda x: rsn.engine.compress(x, ["lzss", "huffman"]),
                              lambda x: rsn.engine.decompress(x, ["lzss", "huffman"]))):
        got = rsn.engine.batch(files, algos, True, workers=workers)
        for f, g in zip(files, got):
            want = single(comp, f)
            assert g == want, (algos, "compress", len(f), seed, rounds)
        streams = [corrupt(g) if g is not None else gen() for g in got]
        back = rsn.engine.batch(streams, algos, False, workers=workers)
        for st, b in zip(streams, back):
            want = single(dec, st)
            assert b == want, (algos, "decompress", len(st), seed, rounds)
        files_checked += 2 * len(files)
    rounds += 1
print(f"fuzz_batch: {rounds} rounds, {files_checked} file operations identical to the single-stream calls (seed {seed})")
