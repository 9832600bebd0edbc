"""GPU parity tests for the engine layering (-algorithm=lzss,huffman) and the .rsn plumbing."""
import hashlib
import json
import os

import pytest

import cases
from raisin_b200 import synth

pytestmark = pytest.mark.gpu
GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "vectors.json")))
ALGOS = ["lzss", "huffman"]


def sha(b):
    return hashlib.sha256(b).hexdigest()


@pytest.mark.parametrize("name", sorted(GOLDEN["layered"]))
def test_layered_cases(rsn, oracle, name):
    data = cases.lzss_cases()[name]
    g = GOLDEN["layered"][name]
    want = oracle.huff_compress(oracle.lzss_compress_async(data, 4096))
    got = rsn.engine.compress(data, ALGOS)
    assert got == want
    assert rsn.engine.compress_fused(data, ALGOS) == want
    assert g["total_len"] == len(got) and g["payload"]["sha256"] == sha(oracle.huff_split(got)[1])
    if "error" in g["roundtrip"]:
        with pytest.raises(rsn.RaisinPanic):
            rsn.engine.decompress(got, ALGOS)
    else:
        back = rsn.engine.decompress(got, ALGOS)
        assert g["roundtrip"]["sha256"] == sha(back)
        assert rsn.engine.decompress_fused(got, ALGOS) == back


def test_rsn_files(rsn, oracle, tmp_path):
    data = synth.mixed(300000, 31, segment=50000)
    p = tmp_path / "corpus.bin"
    p.write_bytes(data)
    out = rsn.engine.CompressFile(ALGOS, str(p))
    assert out.endswith(".rsn")
    blob = open(out, "rb").read()
    assert blob == oracle.huff_compress(oracle.lzss_compress_async(data, 4096, threads=8))
    back = rsn.engine.DecompressFile(ALGOS, out, str(tmp_path / "back.bin"))
    assert open(back, "rb").read() == oracle.lzss_decompress(oracle.huff_decompress(blob))


def test_benchmark_file(rsn):
    text = synth.text(200000, 5)
    r = rsn.engine.BenchmarkFile(["lzss"], text)
    assert r.Lossless and not r.Failed and 80 < r.Ratio < 100
    r = rsn.engine.BenchmarkFile(["huffman"], text)
    assert r.Lossless and 50 < r.Ratio < 60
    r = rsn.engine.BenchmarkFile(ALGOS, text, fused=True)
    assert r.Lossless
    r = rsn.engine.BenchmarkFile(ALGOS, b"a<b>c<d" * 3)
    assert not r.Lossless and not r.Failed      # lossy through the rune layer, by design
    r = rsn.engine.BenchmarkFile(["huffman"], b"")
    assert r.Failed                               # the reference panics -> "DNF" row


def test_benchmark_file_on_device_matches_readme_tables(rsn):
    """rsn_benchmark_file (engine.BenchmarkFile with the histograms and the lossless check on the
    device) against the tables the reference's README prints for its two example files
    (README.md:150-170: ratio, ACTUAL ENTROPY, THEORETICAL ENTROPY, LOSSLESS to two decimals), and
    against the host-side mirror on larger inputs."""
    readme = {  # (engine, file) -> (ratio %, actual entropy, theoretical entropy, lossless)
        ("huffman", b"Hello world!\n"): (307.69, 1.08, 2.20, True),
        ("huffman", b"abcabcabcabcabcabcabcabc\n"): (92.00, 1.24, 1.22, True),
        ("lzss", b"Hello world!\n"): (100.00, 2.20, 2.20, True),
    }
    for (engine, data), (ratio, actual, theo, lossless) in readme.items():
        r = rsn.engine.BenchmarkFile([engine], data, fused=True)
        assert not r.Failed and r.Lossless == lossless
        assert f"{r.Ratio:.2f}" == f"{ratio:.2f}"
        assert f"{r.ActualEntropy:.2f}" == f"{actual:.2f}"
        assert f"{r.Entropy:.2f}" == f"{theo:.2f}"
    for data in (synth.mixed(300000, 8, segment=50000), synth.text(100001, 9), b"a<b>c<d" * 3, bytes(range(256)) * 33):
        a = rsn.engine.BenchmarkFile(ALGOS, data, fused=True)
        b = rsn.engine.BenchmarkFile(ALGOS, data, fused=False)
        assert (a.Lossless, a.Failed) == (b.Lossless, b.Failed)
        assert abs(a.Entropy - b.Entropy) < 1e-12
        assert abs(a.ActualEntropy - b.ActualEntropy) < 1e-5   # float32 in the reference's Result
        assert abs(a.Ratio - b.Ratio) < 1e-3
    assert rsn.engine.BenchmarkFile(["huffman"], b"", fused=True).Failed


def test_mixed_corpus_config3_shape(rsn, oracle):
    """BASELINE config 3 shape at a size the oracle finishes quickly."""
    data = synth.mixed(3 << 20, 3)
    want = oracle.huff_compress(oracle.lzss_compress_async(data, 4096, threads=os.cpu_count() or 1))
    got = rsn.engine.compress_fused(data, ALGOS)
    assert got == want
    assert rsn.engine.decompress_fused(got, ALGOS) == oracle.lzss_decompress(oracle.huff_decompress(want))


def test_batch_of_files(rsn, oracle):
    """BASELINE configs[3] shape (files j = kind j mod 3, seed 1000 + j) at a size the oracle checks quickly."""
    files = [synth.batch_file(j, 48 * 1024) for j in range(24)] + [b"", b"a", b"<\\" * 50]
    got = rsn.engine.batch(files, ALGOS, True, workers=6)
    for f, g in zip(files, got):
        if not f:
            assert g is None          # empty input: huffman layer panics in the reference
        else:
            assert g == oracle.huff_compress(oracle.lzss_compress_async(f, 4096))
    back = rsn.engine.batch([g for g in got if g is not None], ALGOS, False, workers=6)
    k = 0
    for f, g in zip(files, got):
        if g is None:
            continue
        assert back[k] == oracle.lzss_decompress(oracle.huff_decompress(g))
        k += 1
    # lzss alone is lossless on every file
    lz = rsn.engine.batch(files, ["lzss"], True)
    assert rsn.engine.batch(lz, ["lzss"], False) == files


def test_cli_rsn_interchange(rsn, oracle, tmp_path):
    """The C++ host mirror + CLI (raisin_b200/host): `-algorithm=lzss,huffman` writes the same .rsn
    bytes the oracle predicts for stock raisin, decompresses it back, and -benchmark reports Lossless."""
    import subprocess

    cli = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "raisin_b200", "raisin_b200_cli")
    data = synth.text(150000, 9)
    p = tmp_path / "alice.txt"
    p.write_bytes(data)
    subprocess.check_call([cli, "-algorithm=lzss,huffman", str(p)])
    blob = (tmp_path / "alice.txt.rsn").read_bytes()
    assert blob == oracle.huff_compress(oracle.lzss_compress_async(data, 4096, threads=8))
    subprocess.check_call([cli, "-decompress", "-algorithm=lzss,huffman", f"-out={tmp_path / 'back.txt'}",
                           str(tmp_path / "alice.txt.rsn")])
    assert (tmp_path / "back.txt").read_bytes() == data
    out = subprocess.check_output([cli, "-benchmark", "-algorithm=lzss,huffman,[lzss,huffman]", str(p)], text=True)
    rows = [l.split() for l in out.splitlines() if l and not l.startswith(("ENGINE", "File"))]
    assert [r[0] for r in rows] == ["lzss", "huffman", "lzss,huffman"]
    assert all(r[-1] == "true" for r in rows)
    # empty file: the huffman layer panics in the reference -> DNF row, exit code still 0
    e = tmp_path / "empty"
    e.write_bytes(b"")
    out = subprocess.check_output([cli, "-benchmark", "-algorithm=huffman", str(e)], text=True)
    assert "DNF" in out


def test_layer_orders_and_single_layers(rsn, oracle):
    data = synth.text(50000, 13) + b"<tag>\\n" * 20
    for algos in (["lzss"], ["huffman"], ["huffman", "lzss"], ["lzss", "lzss"], ["lzss", "huffman"]):
        want = data
        for a in algos:
            want = oracle.lzss_compress_async(want, 4096) if a == "lzss" else oracle.huff_compress(want)
        got = rsn.engine.compress_fused(data, algos)
        assert got == want, algos
        assert rsn.engine.compress(data, algos) == want, algos
        back = want
        for a in reversed(algos):
            back = oracle.lzss_decompress(back) if a == "lzss" else oracle.huff_decompress(back)
        assert rsn.engine.decompress_fused(got, algos) == back, algos


def test_batch_device_buffers(rsn, oracle):
    """rsn_batch_layers with device-resident inputs and outputs."""
    import ctypes as C

    import torch

    lib = rsn._lib.lib()
    files = [synth.batch_file(j, 40000) for j in range(9)]
    tens = [torch.frombuffer(bytearray(f), dtype=torch.uint8).cuda() for f in files]
    torch.cuda.synchronize()
    n = len(files)
    ins = (C.c_void_p * n)(*[t.data_ptr() for t in tens])
    ns = (C.c_size_t * n)(*[len(f) for f in files])
    outs, out_ns, rcs = (C.c_void_p * n)(), (C.c_size_t * n)(), (C.c_int * n)()
    rsn._lib.check(lib.rsn_batch_layers(b"lzss,huffman", 1, n, ins, ns, outs, out_ns, rcs, 4, 1))
    for i, f in enumerate(files):
        h = (C.c_uint8 * out_ns[i])()
        rsn._lib.check(lib.rsn_dev_download(outs[i], out_ns[i], h, None))
        assert bytes(h) == oracle.huff_compress(oracle.lzss_compress_async(f, 4096))
        lib.rsn_dev_free(outs[i], None)


def test_batch_device_buffers_decompress_stages_header_prefixes(rsn, oracle):
    """Device-resident Huffman decode batches fetch growing prefixes of the streams for the header
    parser (4 KiB, 64 KiB, the rest): headers that end inside each round, a separator that straddles a
    round boundary, alphabets beyond the device tree builder, streams without a separator."""
    import ctypes as C

    import torch

    lib = rsn._lib.lib()
    plain = [synth.batch_file(0, 50000),                      # text: header of a few hundred bytes
             synth.batch_file(2, 200000),                     # random bytes: ~2 800 records, second round
             "".join(chr(0x800 + (i * 104729) % 40000) for i in range(30000)).encode(),  # > 64 KiB of header
             "".join(chr(0x100 + (i * 7919) % 5000) for i in range(6000)).encode(),
             b"abc" * 3000]
    streams = [oracle.huff_compress(f) for f in plain]
    # a header whose 5C 0A sits on the 4096-byte boundary of the first round: pad the alphabet until it does
    for extra in range(200, 1400):
        f = "".join(chr(0x100 + i) for i in range(extra)).encode() + b"x" * 50
        st = oracle.huff_compress(f)
        if st.find(b"\\\n") in (4094, 4095, 4096):
            plain.append(f)
            streams.append(st)
    assert len(streams) > 5
    streams += [b"no separator in here at all" * 400, b"12|a3|b\\\n", b""]  # bad streams fail alone
    want = rsn.engine.batch(streams, ["huffman"], False, workers=2)  # host-buffer path
    for f, w in zip(plain, want):
        try:
            ref = oracle.huff_decompress(oracle.huff_compress(f), strict=False)
        except Exception:
            ref = None
        assert w == ref
    tens = [torch.frombuffer(bytearray(st if st else b"\0"), dtype=torch.uint8).cuda() for st in streams]
    torch.cuda.synchronize()
    n = len(streams)
    ins = (C.c_void_p * n)(*[t.data_ptr() for t in tens])
    ns = (C.c_size_t * n)(*[len(st) for st in streams])
    outs, out_ns, rcs = (C.c_void_p * n)(), (C.c_size_t * n)(), (C.c_int * n)()
    lib.rsn_batch_layers(b"huffman", 0, n, ins, ns, outs, out_ns, rcs, 3, 1)
    for i in range(n):
        if want[i] is None:
            assert rcs[i] != 0, i
            continue
        assert rcs[i] == 0, (i, rcs[i])
        h = (C.c_uint8 * max(1, out_ns[i]))()
        rsn._lib.check(lib.rsn_dev_download(outs[i], out_ns[i], h, None))
        assert bytes(h)[:out_ns[i]] == want[i], i
        lib.rsn_dev_free(outs[i], None)


def test_concurrent_callers(rsn, oracle):
    """engine.BenchmarkSuite calls the codecs from several goroutines at once (engine.go:235-244):
    the C ABI must be re-entrant from multiple OS threads (ctypes releases the GIL during calls)."""
    import threading

    inputs = [synth.mixed(120000 + 7919 * k, 50 + k, segment=30000) for k in range(6)]
    want = [oracle.huff_compress(oracle.lzss_compress_async(d, 4096)) for d in inputs]
    errors = []

    def worker(k):
        try:
            for _ in range(4):
                got = rsn.engine.compress(inputs[k], ALGOS)
                assert got == want[k]
                back = rsn.engine.decompress(got, ALGOS)
                assert back == oracle.lzss_decompress(oracle.huff_decompress(want[k]))
        except Exception as e:  # noqa: BLE001
            errors.append((k, repr(e)))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(len(inputs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
