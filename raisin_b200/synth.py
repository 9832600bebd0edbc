"""Deterministic synthetic corpora for the BASELINE configs.

One generator, seeded, numpy only — the CPU oracle and the GPU path always see identical
bytes.  Kinds: "text" (English-like ASCII without '<', '\\' or bytes >= 0x80), "logs"
(repetitive templated log lines, ~2 % of them carrying '<', '>' and '\\' to exercise the
escape layer), "random" (uniform bytes), "mixed" (1 MiB segments cycling text/logs/random in
40/40/20 proportion).
"""
from __future__ import annotations

import numpy as np

_LETTERS = np.frombuffer(b"etaoinshrdlcumwfgypbvkjxqz", dtype=np.uint8)
_LETTER_P = np.array([12.7, 9.1, 8.2, 7.5, 7.0, 6.7, 6.3, 6.1, 6.0, 4.3, 4.0, 2.8, 2.8, 2.4, 2.4, 2.2, 2.0, 2.0,
                      1.9, 1.5, 1.0, 0.8, 0.15, 0.15, 0.10, 0.07])
_LETTER_CDF = np.cumsum(_LETTER_P / _LETTER_P.sum())


def _vocab(rng: np.random.Generator, words: int = 4096):
    lens = rng.integers(2, 11, size=words)
    offs = np.concatenate(([0], np.cumsum(lens)))
    chars = _LETTERS[np.minimum(np.searchsorted(_LETTER_CDF, rng.random(int(offs[-1]))), 25)]
    return chars, offs, lens


def text(n: int, seed: int = 1) -> bytes:
    """English-like text: 4096-word Zipf(1.0) vocabulary, sentences of 5-20 words, LF ~ every 80 chars."""
    if n <= 0:
        return b""
    rng = np.random.default_rng(seed)
    chars, offs, lens = _vocab(rng)
    words = len(lens)
    zipf = 1.0 / np.arange(1, words + 1)
    cdf = np.cumsum(zipf / zipf.sum())
    avg = float((lens * zipf).sum() / zipf.sum()) + 1.2
    out = []
    have = 0
    while have < n:
        m = int((n - have) / avg * 1.1) + 64
        wi = np.minimum(np.searchsorted(cdf, rng.random(m)), words - 1)
        wl = lens[wi]
        # sentence structure: a sentence ends after 5..20 words
        slen = rng.integers(5, 21, size=m // 5 + 2)
        ends = np.cumsum(slen) - 1
        ends = ends[ends < m]
        is_end = np.zeros(m, dtype=bool)
        is_end[ends] = True
        is_start = np.zeros(m, dtype=bool)
        is_start[0] = True
        is_start[ends[ends + 1 < m] + 1] = True
        comma = (rng.random(m) < 0.05) & ~is_end
        number = rng.random(m) < 0.01
        # trailing bytes per word: "," + " " | ". " / "? " / "! " | " "
        tail = np.where(is_end, 2, np.where(comma, 2, 1))
        tot = wl + tail
        pos = np.concatenate(([0], np.cumsum(tot)))
        buf = np.empty(int(pos[-1]), dtype=np.uint8)
        # word characters
        within = np.arange(int(wl.sum())) - np.repeat(np.concatenate(([0], np.cumsum(wl)))[:-1], wl)
        idx = np.repeat(offs[wi], wl) + within
        dst = np.repeat(pos[:-1], wl) + within
        buf[dst] = chars[idx]
        # numbers: overwrite the word's letters with digits
        if number.any():
            nd = np.repeat(number, wl)
            buf[dst[nd]] = (rng.integers(0, 10, size=int(nd.sum())) + 0x30).astype(np.uint8)
        # capitalise sentence starts (letters only)
        first = pos[:-1][is_start & ~number]
        buf[first] = buf[first] - 32
        # punctuation / separators
        wend = pos[:-1] + wl
        punct = np.frombuffer(b"..?!", dtype=np.uint8)[rng.integers(0, 4, size=m)]
        buf[wend[is_end]] = punct[is_end]
        buf[wend[comma]] = 0x2C
        sp = wend + tail - 1
        buf[sp] = 0x20
        # line breaks: the first separator past each multiple of ~80 columns becomes LF
        col = sp // 80
        brk = np.concatenate(([False], col[1:] != col[:-1]))
        buf[sp[brk]] = 0x0A
        out.append(buf)
        have += buf.size
    return np.concatenate(out)[:n].tobytes()


_TEMPLATES = [
    'level=info msg="request served" route=/api/v1/items status=200',
    'level=info msg="cache hit" key=user:profile shard=3',
    'level=warn msg="slow query" table=orders ms=',
    'level=info msg="connection accepted" peer=10.0.0.',
    'level=error msg="upstream timeout" service=billing retry=',
    'level=debug msg="gc cycle finished" heap_mb=',
    'level=info msg="job scheduled" queue=default worker=',
    'level=info msg="healthcheck ok" component=storage',
]


def logs(n: int, seed: int = 1) -> bytes:
    """Repetitive log lines: ~40 templates, incrementing id, ~2 % with '<', '>' and '\\'."""
    rng = np.random.default_rng(seed)
    templates = list(_TEMPLATES)
    while len(templates) < 40:
        b = _TEMPLATES[len(templates) % len(_TEMPLATES)]
        templates.append(b.replace("level=", f"svc=s{len(templates)} level="))
    out = bytearray()
    i = 0
    counter = int(rng.integers(1000, 100000))
    while len(out) < n:
        m = 4096
        ti = rng.integers(0, len(templates), size=m)
        extra = rng.integers(0, 1000, size=m)
        special = rng.random(m) < 0.02
        for k in range(m):
            t = i + k
            line = "2026-01-01T%02d:%02d:%02dZ %s%d id=%d" % ((t // 3600) % 24, (t // 60) % 60, t % 60,
                                                              templates[ti[k]], extra[k], counter)
            if special[k]:
                line += ' payload=<tag attr="a\\b">'
            counter += 1
            out += line.encode()
            out.append(0x0A)
            if len(out) >= n:
                break
        i += m
    return bytes(out[:n])


def random_bytes(n: int, seed: int = 1) -> bytes:
    return np.random.default_rng(seed).integers(0, 256, size=n, dtype=np.uint8).tobytes()


def mixed(n: int, seed: int = 3, segment: int = 1 << 20) -> bytes:
    """1 MiB segments cycling text / logs / text / logs / random (40/40/20 %)."""
    kinds = ["text", "logs", "text", "logs", "random"]
    out = bytearray()
    k = 0
    while len(out) < n:
        m = min(segment, n - len(out))
        out += generate(kinds[k % len(kinds)], m, seed * 1000 + k)
        k += 1
    return bytes(out)


def generate(kind: str, n: int, seed: int = 1) -> bytes:
    if kind == "text":
        return text(n, seed)
    if kind == "logs":
        return logs(n, seed)
    if kind == "random":
        return random_bytes(n, seed)
    if kind == "mixed":
        return mixed(n, seed)
    if kind == "repetitive":
        return repetitive(n, seed, motif=3000)
    if kind == "uniform8":  # eight equiprobable symbols: 3-bit codes, a Huffman stream that never self-synchronises
        return np.random.default_rng(seed).integers(0, 8, size=n, dtype=np.uint8).tobytes().translate(
            bytes.maketrans(bytes(range(8)), b"abcdefgh"))
    raise ValueError(kind)


def batch_file(j: int, size: int = 262144) -> bytes:
    """File j of BASELINE config 4: kind = j mod 3, seed 1000 + j."""
    return generate(["text", "logs", "random"][j % 3], size, 1000 + j)


def repetitive(n: int, seed: int = 5, motif: int = 2048) -> bytes:
    """BASELINE config 5 shape: a text motif repeated, one byte mutated per repeat."""
    rng = np.random.default_rng(seed)
    m = np.frombuffer(text(motif, seed), dtype=np.uint8)
    reps = n // motif + 1
    buf = np.tile(m, reps)[:n].copy()
    where = (np.arange(reps - 1) * motif + rng.integers(0, motif, size=reps - 1))
    where = where[where < n]
    buf[where] = (rng.integers(0x61, 0x7B, size=where.size)).astype(np.uint8)
    return buf.tobytes()


def config5_into(buf: "np.ndarray", seed: int = 5, motif_len: int = 3000) -> None:
    """BASELINE config 5 stream written into `buf` (uint8): a text motif repeated, one byte mutated
    about every 64 KiB and 1 KiB of fresh text laid over the stream once per MiB, so that L saturates
    for almost every position.  SURVEY 8(d) proposes a 64 KiB motif, but a repeat further back than
    the 4096-byte search buffer can never be referenced (lzss.go:123-129) and such a stream is not
    repetitive to this codec at all (8 MiB of it compress to 91 %); the motif is 3000 bytes instead."""
    n = buf.size
    motif = np.frombuffer(text(motif_len, seed), dtype=np.uint8)
    full = (n // motif_len) * motif_len
    if full:
        buf[:full].reshape(-1, motif_len)[:] = motif
    if n > full:
        buf[full:] = motif[: n - full]
    rng = np.random.default_rng(seed)
    spans = -(-n // 65536)
    where = np.arange(spans, dtype=np.int64) * 65536 + rng.integers(0, 65536, size=spans)
    where = where[where < n]
    buf[where] = rng.integers(0x61, 0x7B, size=where.size).astype(np.uint8)
    mib = n >> 20
    if mib:
        fresh = np.frombuffer(text(min(mib, 4096) * 1024, seed + 1), dtype=np.uint8)
        for k in range(mib):
            at = (k << 20) + 300 + 37 * (k % 1000)
            src = (k % 4096) * 1024
            if at + 1024 <= n:
                buf[at:at + 1024] = fresh[src:src + 1024]


def config5(n: int, seed: int = 5) -> bytes:
    out = np.empty(n, dtype=np.uint8)
    config5_into(out, seed)
    return out.tobytes()
