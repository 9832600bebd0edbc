#!/usr/bin/env python
"""BASELINE configs[4] through the C ABI: rsn_lzss_compress_sharded on ONE process, one host thread
and stream per GPU inside the call (host buffer in, host buffer out).

  python tools/run_sharded_abi.py [MiB] [ngpus] [--check] [--iters K]

Prints one JSON line: wall-clock time of the call (host buffers, so H2D/D2H are inside), bytes that
crossed between GPUs, and with --check: equality with the single-GPU call on the same buffer and a
round trip through the CPU oracle's decoder (the device decoder stops at 4 GiB).
"""
import ctypes as C
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import raisin_b200 as rsn  # noqa: E402
from raisin_b200 import synth  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    mib = float(args[0]) if args else 256
    ngpus = int(args[1]) if len(args) > 1 else 2
    check = "--check" in sys.argv
    iters = int(sys.argv[sys.argv.index("--iters") + 1]) if "--iters" in sys.argv else 3
    n = int(mib * (1 << 20))
    lib = rsn._lib.lib()
    rsn._lib.check(lib.rsn_init(0))
    h_in = lib.rsn_host_alloc(n)  # pinned
    buf = np.frombuffer((C.c_uint8 * n).from_address(h_in), dtype=np.uint8)
    t0 = time.perf_counter()
    synth.config5_into(buf, 5)
    gen_s = time.perf_counter() - t0

    def call(fn, *extra):
        o, on = C.POINTER(C.c_uint8)(), C.c_size_t()
        t0 = time.perf_counter()
        rsn._lib.check(fn(h_in, n, 4096, 0, *extra, C.byref(o), C.byref(on)))
        return o, on.value, (time.perf_counter() - t0) * 1e3

    times = []
    digest = None
    for it in range(iters):
        o, on, ms = call(lib.rsn_lzss_compress_sharded, ngpus)
        times.append(ms)
        if it == iters - 1:
            comp = rsn._lib.bytes_at(o, on)
            digest = hashlib.sha256(comp).hexdigest()
        lib.rsn_free(o)
    best = min(times)
    line = {"workload": "config 5: one repetitive stream, lzss compress (variant A, W 4096), match search sharded by "
                        "position range, rsn_lzss_compress_sharded (host buffers in and out)",
            "n_gpus": ngpus, "stream_bytes": n, "ms_per_call": best, "all_ms": [round(t, 1) for t in times],
            "GBps": n / (best * 1e-3) / 1e9, "compressed_bytes": on, "sha256": digest,
            "gpu_to_gpu_bytes": int(lib.rsn_sharded_peer_bytes()),
            "per_position_arrays_gathered": 0, "input_generated_s": round(gen_s, 1)}
    if check:
        o1, on1, ms1 = call(lib.rsn_lzss_compress)
        single = rsn._lib.bytes_at(o1, on1)
        lib.rsn_free(o1)
        line["single_gpu_ms"] = ms1
        line["identical_to_single_gpu"] = single == comp
        from oracle import pyoracle as po

        t0 = time.perf_counter()
        back = po.lzss_decompress(comp)
        line["oracle_decode_s"] = round(time.perf_counter() - t0, 1)
        line["roundtrip_ok"] = len(back) == n and hashlib.sha256(back).digest() == hashlib.sha256(buf).digest()
    print(json.dumps(line), flush=True)
    lib.rsn_host_free(h_in)


if __name__ == "__main__":
    main()
