#!/usr/bin/env python
"""Write the named test inputs (tests/cases.py) to tests/golden/inputs.json.gz so that a machine with a
Go toolchain can run the stock reference on exactly these bytes (go/golden/golden_dump_test.go) and
turn the oracle-derived vectors into true goldens.  Run from the repo root:
    python tests/tools/export_inputs.py"""
import gzip
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402

out = {"lzss": {k: v.hex() for k, v in cases.lzss_cases().items()},
       "huffman": {k: v.hex() for k, v in cases.huffman_cases().items()}}
path = os.path.join(ROOT, "tests", "golden", "inputs.json.gz")
with gzip.GzipFile(path, "wb", mtime=0) as fh:
    fh.write(json.dumps(out, sort_keys=True).encode())
print("wrote", path, {k: len(v) for k, v in out.items()}, os.path.getsize(path), "bytes")
