#!/usr/bin/env python
"""Regenerate tests/golden/vectors.json from the CPU oracle (oracle/raisin_oracle.c).

Run from the repo root:  python tests/tools/make_golden.py
The fixtures pin the oracle's outputs on the named cases of tests/cases.py so that (a) the
oracle cannot drift silently and (b) the GPU parity tests have expectations that do not
depend on the oracle library being rebuilt identically.  Small outputs are stored as hex,
large ones as length + sha256.
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def rec(b: bytes):
    d = {"len": len(b), "sha256": hashlib.sha256(b).hexdigest()}
    if len(b) <= 96:
        d["hex"] = b.hex()
    return d


def attempt(fn, *a):
    try:
        return rec(fn(*a))
    except po.OracleError as e:
        return {"error": e.name}


def main():
    out = {"lzss": {}, "huffman": {}, "layered": {}}
    for name, data in cases.lzss_cases().items():
        comp = po.lzss_compress_async(data, 4096)
        e = {"input": rec(data), "async_w4096": rec(comp), "iter_w4096": rec(po.lzss_compress_iter(data, 4096)),
             "async_w1024": rec(po.lzss_compress_async(data, 1024)),
             "decompress_async": attempt(po.lzss_decompress, comp),
             "decompress_iter": attempt(po.lzss_decompress, po.lzss_compress_iter(data, 4096))}
        out["lzss"][name] = e
        if data:
            lay = po.huff_compress(comp) if comp else None
            if lay is not None:
                hd, pl = po.huff_split(lay)
                back = attempt(po.huff_decompress, lay)
                e2 = {"payload": rec(pl), "header_len": len(hd), "total_len": len(lay), "huff_decompress": back}
                try:
                    e2["roundtrip"] = rec(po.lzss_decompress(po.huff_decompress(lay)))
                except po.OracleError as ex:
                    e2["roundtrip"] = {"error": ex.name}
                out["layered"][name] = e2
    for name, data in cases.huffman_cases().items():
        comp = po.huff_compress(data)
        hd, pl = po.huff_split(comp)
        hm = po.huff_header_map(hd) if not name.startswith("backslash_only") else {}
        out["huffman"][name] = {
            "input": rec(data), "compressed": rec(comp), "payload": rec(pl), "header_len": len(hd),
            "header_symbols": len(hm),
            "decompress": attempt(po.huff_decompress, comp),
            "decompress_strict": attempt(po.huff_decompress, comp, True),
        }
    path = os.path.join(ROOT, "tests", "golden", "vectors.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    print("wrote", path, {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
