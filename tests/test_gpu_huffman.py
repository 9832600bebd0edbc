"""GPU parity tests for the Huffman path (C ABI vs oracle vs golden).  Payload bytes are compared
byte for byte; the header is compared as the rune->freq map the reference's decodeTree reads,
and additionally byte for byte against the oracle's canonical record order."""
import hashlib
import json
import os

import numpy as np
import pytest

import cases
from raisin_b200 import synth

pytestmark = pytest.mark.gpu
GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "vectors.json")))
ERR = {"empty_input": -10, "no_separator": -11, "bad_header": -12, "truncated": -13, "guard": -14,
       "single_leaf_loop": -16}


def sha(b):
    return hashlib.sha256(b).hexdigest()


@pytest.mark.parametrize("name", sorted(GOLDEN["huffman"]))
def test_compress_cases(rsn, oracle, name):
    data = cases.huffman_cases()[name]
    g = GOLDEN["huffman"][name]
    got = rsn.huffman.Compress(data)
    want = oracle.huff_compress(data)
    hd, pl = oracle.huff_split(got)
    whd, wpl = oracle.huff_split(want)
    assert pl == wpl
    assert g["payload"]["sha256"] == sha(pl)
    if name != "backslash_only":
        assert oracle.huff_header_map(hd) == oracle.huff_header_map(whd)
    assert got == want
    for key, strict in (("decompress", False), ("decompress_strict", True)):
        if "error" in g[key]:
            with pytest.raises(rsn.RaisinPanic) as ei:
                rsn.huffman.Decompress(got, strict)
            assert ei.value.rc == ERR[g[key]["error"]]
        else:
            back = rsn.huffman.Decompress(got, strict)
            assert g[key]["sha256"] == sha(back)
            assert back == oracle.huff_decompress(got, strict)


def test_empty_input_panics(rsn):
    with pytest.raises(rsn.RaisinPanic) as ei:
        rsn.huffman.Compress(b"")
    assert ei.value.rc == -10


def test_decoder_failure_modes(rsn, oracle):
    good = oracle.huff_compress(b"abracadabra")
    hd, pl = oracle.huff_split(good)
    for blob, err in [(b"no separator here", "no_separator"), (hd + b"\\\n", "truncated"),
                      (hd + b"\\\n" + bytes([9]) + pl[1:2], "truncated"), (b"5|\\\n\x00", "bad_header"),
                      (b"\\\n\x00", "bad_header"), (b"3|a\\\n\x00\xff", "single_leaf_loop"),
                      (good[:-1], "truncated")]:
        with pytest.raises(rsn.RaisinPanic) as ei:
            rsn.huffman.Decompress(blob)
        assert ei.value.rc == ERR[err], blob


def test_decode_headers_in_any_order(rsn, oracle):
    """Go writes header records in map order; the decoder must accept any order."""
    data = cases.huffman_cases()["text_8k"]
    comp = oracle.huff_compress(data)
    hd, pl = oracle.huff_split(comp)
    recs = []
    i = 0
    m = oracle.huff_header_map(hd)
    from oracle import go_literal as gl
    for r in sorted(m, key=lambda r: (m[r] * 7919 + r) % 1013):
        recs.append(str(m[r]).encode() + b"|" + (b"\\n" if r == 10 else gl.go_string_rune(r)))
    if recs[-1].endswith(b"|\\"):
        recs[0], recs[-1] = recs[-1], recs[0]
    shuffled = b"".join(recs) + b"\\\n" + pl
    assert rsn.huffman.Decompress(shuffled) == data == oracle.huff_decompress(shuffled)


def test_random_small_inputs(rsn, oracle):
    rng = np.random.default_rng(777)
    for trial in range(100):
        n = int(rng.integers(1, 900))
        alpha = [b"ab", b"abc|\\\n", "aé☃😀".encode(), bytes(range(256)), b"the quick brown fox "][trial % 5]
        data = bytes(alpha[i] for i in rng.integers(0, len(alpha), size=n))
        got = rsn.huffman.Compress(data)
        assert got == oracle.huff_compress(data), (trial, n)
        try:
            want = oracle.huff_decompress(got)
        except oracle.OracleError as e:
            with pytest.raises(rsn.RaisinPanic) as ei:
                rsn.huffman.Decompress(got)
            assert ei.value.rc == ERR[e.name]
        else:
            assert rsn.huffman.Decompress(got) == want


def test_one_mib_text_config1(rsn, oracle):
    """BASELINE config 1: 1 MiB English-like text.  > 900000 bits, so the stock decoder panics
    (strict), and the guard-lifted decode restores the input."""
    data = synth.text(1 << 20, 1)
    got = rsn.huffman.Compress(data)
    assert got == oracle.huff_compress(data)
    with pytest.raises(rsn.RaisinPanic) as ei:
        rsn.huffman.Decompress(got, strict_limits=True)
    assert ei.value.rc == -14
    assert rsn.huffman.Decompress(got) == data


def test_large_alphabet_random(rsn, oracle):
    data = synth.random_bytes(4 << 20, 5)
    got = rsn.huffman.Compress(data)
    assert got == oracle.huff_compress(data)
    back = rsn.huffman.Decompress(got)
    assert back == oracle.huff_decompress(got)
    assert back != data  # invalid UTF-8 became U+FFFD: lossy by design of the reference


def test_full_size_roundtrip(rsn):
    data = synth.text(64 << 20, 2)
    comp = rsn.huffman.Compress(data)
    assert rsn.huffman.Decompress(comp) == data


def test_no_state_between_calls(rsn):
    """The reference never resets `answer` in Decompress (huffman.go:129); we deliberately do."""
    a = rsn.huffman.Compress(b"first message")
    b = rsn.huffman.Compress(b"second one")
    assert rsn.huffman.Decompress(a) == b"first message"
    assert rsn.huffman.Decompress(b) == b"second one"


def test_non_synchronising_stream(rsn, oracle):
    """Eight equiprobable symbols give eight 3-bit codes: a decoder that starts inside a code never
    re-aligns (256 is not a multiple of 3), so the speculative fix-up would move one subsequence per
    round.  The decoder must notice after a bounded number of rounds and take the exact path
    (transfer functions over the candidate starts); same bytes as the oracle, single call and batch."""
    data = synth.generate("uniform8", 4 << 20, 77)
    comp = rsn.huffman.Compress(data)
    assert comp == oracle.huff_compress(data)
    assert rsn.huffman.Decompress(comp) == oracle.huff_decompress(comp) == data
    small = [synth.generate("uniform8", 100000 + 977 * k, 80 + k) for k in range(5)] + [synth.text(90000, 3)]
    blobs = [oracle.huff_compress(x) for x in small]
    assert rsn.engine.batch(blobs, ["huffman"], False, workers=2) == small
    # sixteen symbols: 4-bit codes DO re-align with 256-bit subsequences (control case)
    data16 = bytes(0x61 + (b & 15) for b in synth.random_bytes(1 << 20, 5))
    assert rsn.huffman.Decompress(rsn.huffman.Compress(data16)) == data16
