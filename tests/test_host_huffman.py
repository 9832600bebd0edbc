"""CPU tests of the product's host-side Huffman logic (raisin_b200/csrc/huff_host.cpp: leaf order,
Go's container/heap replay, codes, header formatting and decodeTree's parse) against the oracle.
The sources are compiled into a small harness (tests/host/huff_host_check.cpp) with nvcc in host
mode; nothing here needs a GPU."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import go_literal
from oracle import pyoracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "raisin_b200", "csrc")
BUILD = os.path.join(ROOT, "tests", "_build")


@pytest.fixture(scope="module")
def harness():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "huff_host_check")
    srcs = [os.path.join(ROOT, "tests", "host", "huff_host_check.cpp"), os.path.join(CSRC, "huff_host.cpp")]
    if not os.path.exists(exe) or any(os.path.getmtime(s) > os.path.getmtime(exe) for s in srcs):
        subprocess.run([nvcc, "-O2", "-std=c++17", "-x", "cu", "-I", CSRC, *srcs, "-o", exe], check=True,
                       capture_output=True, timeout=600)
    return exe


def run(exe, mode, text):
    return subprocess.run([exe, mode], input=text.encode(), capture_output=True, check=True, timeout=120).stdout.decode()


def alphabets():
    rng = np.random.default_rng(7)
    yield [65], [5]                                      # single leaf: empty code
    yield [65, 66], [1, 1]
    yield list(range(97, 123)), [1] * 26                 # all tied
    for k, hi in ((3, 3), (17, 2), (200, 4), (1000, 3), (3000, 2), (500, 1000), (64, 10 ** 6), (2800, 6)):
        runes = rng.choice(0x3000, size=k, replace=False) + 1
        runes = [int(r) for r in runes if not 0xD800 <= r <= 0xDFFF]
        yield runes, [int(f) for f in rng.integers(1, hi + 1, size=len(runes))]
    # Fibonacci-like frequencies: a deep tree
    fib = [1, 1]
    while len(fib) < 40:
        fib.append(fib[-1] + fib[-2])
    yield list(range(200, 240)), fib


def test_tree_and_codes_match_oracle(harness):
    for runes, freqs in alphabets():
        text = f"{len(runes)}\n" + "".join(f"{r} {f}\n" for r, f in zip(runes, freqs))
        lines = run(harness, "codes", text).splitlines()
        got = {}
        for ln in lines[:-1]:
            r, l, c = ln.split()
            got[int(r)] = (int(l), int(c, 16))
        code, ln = po.huff_code_table(runes, freqs)
        want = {int(r): (int(l), int(c)) for r, l, c in zip(runes, ln, code)}
        assert got == want, (len(runes), max(freqs))


def test_header_matches_oracle(harness):
    """huff_header against the oracle's own Compress output (canonical rune order, LF as \\n)."""
    rng = np.random.default_rng(11)
    for k in (1, 2, 5, 60, 300):
        runes = [10, 0x5C, 0x7C, 0x31][: min(k, 4)] + [int(r) for r in rng.choice(0x2000, size=k, replace=False) + 0x80
                                                       if not 0xD800 <= r + 0 <= 0xDFFF]
        runes = list(dict.fromkeys(runes))
        freqs = [int(f) for f in rng.integers(1, 40, size=len(runes))]
        data = "".join(chr(r) * f for r, f in zip(runes, freqs)).encode()
        want_hdr, _ = po.huff_split(po.huff_compress(data))
        text = f"{len(runes)}\n" + "".join(f"{r} {f}\n" for r, f in zip(runes, freqs))
        got_hdr = bytes.fromhex(run(harness, "codes", text).splitlines()[-1].split()[1])
        assert got_hdr == want_hdr, k
        # the same leaves handed over in ascending rune order (what the batch path does: no sort inside)
        pairs = sorted(zip(runes, freqs))
        text = f"{len(pairs)}\n" + "".join(f"{r} {f}\n" for r, f in pairs)
        got_hdr = bytes.fromhex(run(harness, "codes", text).splitlines()[-1].split()[1])
        assert got_hdr == want_hdr, k


def test_header_parse_matches_decode_tree(harness):
    """huff_parse_header against the literal transcription of decodeTree (huffman.go:196-227):
    duplicates (last assignment wins), the \\n escape, stray bytes, records cut short."""
    rng = np.random.default_rng(13)
    headers = [b"3|a2|b", b"3|a2|a", b"12|\\n1|x", b"1|\\\\2|\\n", b"x9y|z", b"|a", b"5|", b"5|\\", b"7|\xc3\xa92|\xe2\x98\x83",
               b"1|a1|b1|c1|a9|b", b"", b"123", b"4|\xff3|\x80"]
    alpha = b"0123456789|\\nab\xc3\xa9\xff "
    for _ in range(60):
        headers.append(bytes(alpha[i] for i in rng.integers(0, len(alpha), size=int(rng.integers(1, 60)))))
    big = "".join(f"{int(f)}|{chr(int(r))}" for r, f in zip(rng.choice(0x2000, 3000, replace=False) + 0x100,
                                                             rng.integers(1, 9, 3000))).encode()
    headers.append(big)
    # more than 2^16 records: the dense-table path, with repeated runes (the last assignment wins)
    pool = [int(r) for r in rng.choice(0x20000, 40000, replace=False) + 0x100 if not 0xD800 <= r <= 0xDFFF]
    picks = rng.integers(0, len(pool), 70000)
    headers.append("".join(f"{int(f)}|{chr(pool[int(i)])}" for i, f in zip(picks, rng.integers(1, 99, 70000))).encode())
    for h in headers:
        out = run(harness, "parse", (h.hex() or "-") + "\n").splitlines()
        try:
            want = go_literal.decodeTree(h)
        except Exception:
            want = None
        if want is None:
            assert out[0] == "bad", h
            continue
        assert out[0].startswith("ok"), h
        got = {int(a): int(b) for a, b in (ln.split() for ln in out[1:])}
        assert got == {int(k): int(v) for k, v in want.items()}, h
