#!/usr/bin/env python
"""BASELINE configs[2]: layered lzss,huffman on a large synthetic mixed corpus (text + repetitive
logs + random bytes) on one B200, checked byte for byte against the CPU oracle, with timings.
usage: python tests/tools/validate_large.py [MiB=1024]"""
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

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = mib << 20
t0 = time.time()
data = synth.mixed(n, 3)
gen_s = time.time() - t0
algos = ["lzss", "huffman"]
rsn.engine.compress_fused(data[: 1 << 20], algos)  # warm-up
t0 = time.perf_counter()
comp = rsn.engine.compress_fused(data, algos)
t_c = time.perf_counter() - t0
t0 = time.perf_counter()
back = rsn.engine.decompress_fused(comp, algos)
t_d = time.perf_counter() - t0
cores = os.cpu_count() or 1
t0 = time.time()
want_lz = po.lzss_compress_async(data, 4096, threads=cores)
want = po.huff_compress(want_lz)
t_oc = time.time() - t0
t0 = time.time()
want_back = po.lzss_decompress(po.huff_decompress(want))
t_od = time.time() - t0
print(json.dumps({
    "workload": f"layered lzss,huffman, {mib} MiB synthetic mixed corpus (BASELINE configs[2])",
    "input_bytes": n, "compressed_bytes": len(comp), "ratio_pct": 100.0 * len(comp) / n,
    "gpu_compress_s_host_buffers": t_c, "gpu_decompress_s_host_buffers": t_d,
    "gpu_compress_GBps": n / t_c / 1e9, "gpu_decompress_GBps": n / t_d / 1e9,
    "compress_identical_to_oracle": comp == want, "decompress_identical_to_oracle": back == want_back,
    "lossless": back == data, "note": "lossless=false is the reference's own behaviour on non-UTF-8 bytes (SURVEY F6)",
    "oracle_compress_s": t_oc, "oracle_decompress_s": t_od, "oracle_threads": cores,
    "oracle_mode": "bisection mode (same results as the literal mode, fewer searches)",
    "sha256_compressed": hashlib.sha256(comp).hexdigest(), "generate_s": gen_s,
}))
