"""Named, deterministic inputs shared by the golden generator and the parity tests
(the adversarial classes of SURVEY A.9 plus the synthetic corpora of 8(d))."""
from __future__ import annotations

import numpy as np

from raisin_b200 import synth


def _rng(seed):
    return np.random.default_rng(seed)


def lzss_cases():
    """name -> bytes.  Small enough for the oracle to finish in well under a second each."""
    c = {}
    c["empty"] = b""
    c["one_byte"] = b"a"
    c["hello"] = b"Hello world!\n"
    c["abc8"] = b"abcabcabcabcabcabcabcabc\n"
    c["a20"] = b"a" * 20
    c["a40"] = b"a" * 40
    c["ab10"] = b"ab" * 10
    c["digits3"] = b"0123456789" * 3
    c["abcdefgh4"] = b"abcdefgh" * 4 + b"!"
    c["escapes"] = bytes.fromhex("783c795c7aff77")
    c["lt_only"] = b"<" * 50
    c["backslashes"] = b"\\" * 33 + b"x" + b"\\" * 32 + b"<" + b"\\\\\xff\\<"
    c["ff_run"] = b"\xff" * 40 + b"\\\xff" * 9
    c["token_innards"] = b"12,34>56,7>>,,1,2,3>" * 6
    c["layer_lossy"] = b"a<b>c<d" * 3
    for p in range(1, 9):
        c[f"period{p}"] = bytes((0x61 + (i % p)) for i in range(300))
    c["a5000"] = b"a" * 5000
    c["a9000"] = b"a" * 9000
    c["ab4500"] = b"ab" * 4500
    # match ends exactly at EOF / window boundaries i == W, W+1
    blk = _rng(7).integers(0x61, 0x7B, size=64, dtype=np.uint8).tobytes()
    c["blk_at_eof"] = blk + _rng(8).integers(0x41, 0x5B, size=200, dtype=np.uint8).tobytes() + blk
    noise = _rng(9).integers(0x41, 0x5B, size=4096 - 64, dtype=np.uint8).tobytes()
    c["win_exact"] = blk + noise + blk            # second copy starts at i == 4096: source at distance W
    c["win_plus1"] = blk + noise + b"Z" + blk     # distance W+1: first byte of the source is out of reach
    c["win_minus1"] = blk + noise[:-1] + blk
    # digit-count thresholds of the emit rule (L = 5, 6, 9, 10, 11, 99, 100)
    parts = []
    r = _rng(10)
    for L in (5, 6, 9, 10, 11, 99, 100, 101, 999, 1000):
        s = r.integers(0x61, 0x7B, size=L, dtype=np.uint8).tobytes()
        parts.append(s + b"#" + s + b"%")
    c["emit_thresholds"] = b"".join(parts)
    c["text_8k"] = synth.text(8192, 11)
    c["text_64k"] = synth.text(65536, 12)
    c["logs_64k"] = synth.logs(65536, 13)
    c["random_16k"] = synth.random_bytes(16384, 14)
    c["mixed_96k"] = synth.mixed(98304, 15, segment=16384)
    c["repetitive_64k"] = synth.repetitive(65536, 16, motif=1024)
    c["binary_lowentropy"] = _rng(17).integers(0, 4, size=20000, dtype=np.uint8).tobytes()
    c["ragged_4097"] = synth.text(4097, 18)
    c["ragged_4095"] = synth.text(4095, 19)
    c["ragged_12289"] = synth.logs(12289, 20)
    return c


def huffman_cases():
    c = {}
    c["one_byte"] = b"a"
    c["aaaa"] = b"aaaa"
    c["ab"] = b"ab"
    c["aab"] = b"aab"
    c["hello"] = b"Hello world!\n"
    c["abc8"] = b"abcabcabcabcabcabcabcabc\n"
    c["invalid_utf8"] = bytes.fromhex("41ff42c328e28241eda08041")
    c["utf8_ok"] = "héé ☃".encode()
    c["utf8_4byte"] = "a😀b😀😀c𝄞".encode() * 5
    c["overlong_surrogates"] = bytes.fromhex("c080e08080f0808080eda080edbfbff4908080f5808080c2") + b"xyz"
    c["truncated_tail"] = "héllo wörld ☃☃".encode() + bytes.fromhex("e298")
    c["digits_pipes"] = b"1|2|3||4\\5\\\\6|\\n\r\n\x00\x00||99" * 7
    c["symbols_special"] = b"|\\\n\r\x00 0123456789" * 11
    c["two_symbols"] = b"ab" * 100 + b"a" * 37
    c["equal_freqs"] = bytes(range(0x30, 0x30 + 64)) * 9
    for k in range(8):
        c[f"bits_mod8_{k}"] = b"ab" * 50 + b"c" * (3 + k)
    c["backslash_only"] = b"\\" * 10          # header would end in '\\': decoder fails
    c["backslash_max"] = b"a\\b\\c"           # '\\' is the largest rune: canonical order swaps the last two
    c["text_8k"] = synth.text(8192, 21)
    c["text_256k"] = synth.text(262144, 22)    # > 900000 bits: strict vs lifted
    c["logs_64k"] = synth.logs(65536, 23)
    c["random_64k"] = synth.random_bytes(65536, 24)   # ~2k distinct runes incl. U+FFFD (lossy by design)
    c["mixed_96k"] = synth.mixed(98304, 25, segment=16384)
    c["dna_like"] = _rng(26).integers(0, 4, size=30000, dtype=np.uint8).astype(np.uint8).tobytes().translate(
        bytes.maketrans(bytes(range(4)), b"ACGT"))
    c["eight_uniform"] = _rng(27).integers(0, 8, size=30000, dtype=np.uint8).tobytes().translate(
        bytes.maketrans(bytes(range(8)), b"abcdefgh"))
    c["skewed"] = (b"e" * 4000 + b"t" * 2000 + b"a" * 1000 + b"o" * 500 + b"i" * 250 + b"n" * 125 + b"s" * 60 +
                   b"h" * 30 + b"r" * 15 + b"d" * 7 + b"l" * 3 + b"u" * 2 + b"z")
    return c


# ---- reference-recorded artefacts ----------------------------------------------------------
# /root/reference/ai/data.json holds results the reference's own engine.BenchmarkFile produced
# with the `lzss` engine of that era (the iterative lz.Compress, window 4096) on files that can
# be rebuilt offline.  Each record: what the Go code measured on its own output, i.e. pins for
# lz.Compress + lz.Decompress that do not come from our restatement.
#   compressed_ratio  = float32(len(compressed)) / float32(len(input)) * 100   (engine.go:409)
#   compressed_entropy = float32(natural-log Shannon entropy of the compressed bytes' histogram)
#   lossless          = DeepEqual(input, Decompress(Compress(input)))           (engine.go:408)

def entropy_nat(b: bytes) -> float:
    if not b:
        return 0.0
    h = np.bincount(np.frombuffer(b, dtype=np.uint8), minlength=256).astype(np.float64)
    p = h[h > 0] / len(b)
    return float(-(p * np.log(p)).sum())


def pi_txt() -> bytes:
    """Canterbury pi.txt (first 10^6 digits of pi), rebuilt by tests/tools/make_pi_fixture.py."""
    import lzma
    import os

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pi_1e6.xz"), "rb") as f:
        return lzma.decompress(f.read())


def reference_recorded():
    """name -> (input bytes, recorded input entropy, compressed_ratio, compressed_entropy, lossless,
    data.json line of the lzss record)."""
    alpha = (b"abcdefghijklmnopqrstuvwxyz" * 3847)[:100000]
    return {
        "aaa.txt": (b"a" * 100000, 0.0, 0.4000000059604645, 2.591623306274414, True, 2016),
        "alphabet.txt": (alpha, 3.2580965336217447, 100.0, 3.25809645652771, True, 2158),
        "a.txt": (b"a", 0.0, 100.0, 0.0, True, 2726),
        "pi.txt": (pi_txt(), 2.3025823330123467, 100.0, 2.3026533126831055, False, 3365),
    }


def recorded_result(data: bytes, comp: bytes, back: bytes):
    """What engine.BenchmarkFile of that era would have recorded for these buffers."""
    ratio = float(np.float32(len(comp)) / np.float32(len(data)) * np.float32(100))
    return ratio, float(np.float32(entropy_nat(comp))), back == data
