#!/usr/bin/env python
"""Compare the stock reference's outputs (go/golden/golden_dump_test.go, run on a machine with Go)
with tests/golden/vectors.json.  LZSS fields are compared byte for byte (length + SHA-256); Huffman
outputs by payload bytes, by the header as a set of `freq|symbol` records (Go's map order is random)
and by total length.   usage: python tests/tools/compare_reference_outputs.py reference_outputs.json"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

ref = json.load(open(sys.argv[1]))
gold = json.load(open(os.path.join(ROOT, "tests", "golden", "vectors.json")))
bad = checked = 0


def same(a, b, what):
    global bad, checked
    checked += 1
    if ("error" in a) != ("error" in b) or ("error" not in a and (a["len"], a["sha256"]) != (b["len"], b["sha256"])):
        bad += 1
        print("MISMATCH", what, a, b)


def split(blob):
    hd, pl = po.huff_split(blob)
    try:
        return po.huff_header_map(hd), pl
    except Exception:  # a header the reference's own decodeTree panics on (e.g. it ends in a backslash record)
        return None, pl


for name, e in ref["vectors"]["lzss"].items():
    g = gold["lzss"][name]
    for k in ("async_w4096", "async_w1024", "iter_w4096", "decompress_async", "decompress_iter"):
        same(e[k], g[k], f"lzss/{name}/{k}")
for key, hx in ref["huffman_full_hex"].items():
    kind, name = key.split("/", 1)
    blob = bytes.fromhex(hx)
    hm, pl = split(blob)
    g = gold[kind][name]
    checked += 1
    want_len = g["compressed"]["len"] if kind == "huffman" else g["total_len"]
    if hashlib.sha256(pl).hexdigest() != g["payload"]["sha256"] or len(blob) != want_len:
        bad += 1
        print("MISMATCH", key, "payload/length")
    if kind == "huffman" and hm is not None and len(hm) != g["header_symbols"]:
        bad += 1
        print("MISMATCH", key, "header symbols")
for name, e in ref["vectors"]["huffman"].items():
    same(e["decompress"], gold["huffman"][name]["decompress_strict"], f"huffman/{name}/decompress")
print(f"{checked} comparisons, {bad} mismatches")
sys.exit(1 if bad else 0)
