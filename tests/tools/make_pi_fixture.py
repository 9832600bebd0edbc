#!/usr/bin/env python
"""Rebuild tests/golden/pi_1e6.xz: the first 1 000 000 decimal digits of pi ("31415926…", no
point, no newline) = the Canterbury "large/misc" corpus file pi.txt that the reference's recorded
benchmark (/root/reference/ai/data.json:3341-3344) ran on.  The file itself is not in the
reference tree; it is recomputed here (Chudnovsky, binary splitting, plain Python ints, ~15 s)
and checked against what data.json records about the original: size 1 000 000 and byte-histogram
entropy 2.3025823330123467 nat (matches to the last printed digit).
"""
import hashlib
import lzma
import math
import os
import sys

import numpy as np

SHA256 = "387877db67fdddbde761c053c4376e0b411b10fd2b126fd8b1249963cb628877"
RECORDED_ENTROPY = 2.3025823330123467  # ai/data.json:3343


def pi_digits(count: int) -> bytes:
    sys.set_int_max_str_digits(0)

    def bs(a, b):
        if b - a == 1:
            if a == 0:
                p = q = 1
            else:
                p = (6 * a - 5) * (2 * a - 1) * (6 * a - 1)
                q = a * a * a * 10939058860032000
            t = p * (13591409 + 545140134 * a)
            return p, q, -t if a & 1 else t
        m = (a + b) // 2
        p1, q1, t1 = bs(a, m)
        p2, q2, t2 = bs(m, b)
        return p1 * p2, q1 * q2, q2 * t1 + p1 * t2

    d = count + 20
    _, q, t = bs(0, d // 14 + 2)
    one = 10 ** d
    pi = (q * 426880 * math.isqrt(10005 * one * one)) // t
    return str(pi)[:count].encode()


def main():
    data = pi_digits(1_000_000)
    assert len(data) == 1_000_000 and data[:10] == b"3141592653"
    assert hashlib.sha256(data).hexdigest() == SHA256
    h = np.bincount(np.frombuffer(data, dtype=np.uint8), minlength=256).astype(np.float64)
    p = h[h > 0] / len(data)
    assert float(-(p * np.log(p)).sum()) == RECORDED_ENTROPY
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "golden", "pi_1e6.xz")
    with open(out, "wb") as f:
        f.write(lzma.compress(data, preset=9 | lzma.PRESET_EXTREME))
    print(out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
