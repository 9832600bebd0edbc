#!/usr/bin/env python
"""BASELINE configs[0]: `raisin -benchmark -algorithm=huffman` on a 1 MiB synthetic English-like text
file.  Times engine.BenchmarkFile's timed region (compress + decompress, wall clock, host buffers,
engine.go:395-408) through the C ABI, checks the compressed bytes and the round trip against the
CPU oracle and prints the benchmark row next to the oracle's own time on one host core (the
reference's Huffman coder is single-threaded).
usage: python tests/tools/run_config1.py [MiB=1]"""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import raisin_b200 as rsn  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from raisin_b200 import synth  # noqa: E402

mib = float(sys.argv[1]) if len(sys.argv) > 1 else 1
n = int(mib * (1 << 20))
data = synth.generate("text", n, 11)
algos = ["huffman"]
for _ in range(3):  # warm-up: context, pools
    rsn.engine.BenchmarkFile(algos, data, fused=True)
runs = []
for _ in range(10):
    t0 = time.perf_counter()
    comp = rsn.huffman.Compress(data)
    back = rsn.huffman.Decompress(comp)
    runs.append(time.perf_counter() - t0)
fused = [rsn.engine.BenchmarkFile(algos, data, fused=True) for _ in range(10)]
row = fused[-1]
t0 = time.perf_counter()
want = po.huff_compress(data)
want_back = po.huff_decompress(want)
t_o = time.perf_counter() - t0
best = min(runs)
print(json.dumps({
    "workload": f"-benchmark -algorithm=huffman, {mib:g} MiB synthetic text (BASELINE configs[0])",
    "input_bytes": n, "compressed_bytes": len(comp), "ratio_pct": row.Ratio,
    "entropy": row.Entropy, "actual_entropy": row.ActualEntropy, "lossless": row.Lossless,
    "gpu_timed_region_ms_host_buffers_best_of_10": best * 1e3,
    "gpu_timed_region_ms_host_buffers_median": sorted(runs)[len(runs) // 2] * 1e3,
    "gpu_benchmark_file_one_call_ms_best_of_10": min(r.Seconds for r in fused) * 1e3,
    "gpu_GBps": 2 * n / best / 1e9,
    "oracle_timed_region_ms_one_core": t_o * 1e3, "oracle_GBps": 2 * n / t_o / 1e9,
    "compress_identical_to_oracle": comp == want, "decompress_identical_to_oracle": back == want_back,
    "sha256_compressed": hashlib.sha256(comp).hexdigest(),
    "note": "GB/s = (bytes in + bytes back) / timed region; headers compared as rune->frequency maps is not needed: "
            "both sides write records in ascending rune order",
}))
