"""GPU test of the position-range sharded LZSS compress (BASELINE configs[4] shape): every shard's
match arrays are computed by the CUDA kernel on a slice with window halo + look-ahead, assembled,
and the merge/emit runs once on the assembled arrays.  Shards are emulated in one process here;
tools/run_sharded.py does the same across ranks with an NCCL all-gather."""
import ctypes as C

import numpy as np
import pytest

from raisin_b200 import parallel, synth

pytestmark = pytest.mark.gpu


def _gpu_match(lib, rsn, slice_t, window, sp):
    import torch

    out = torch.empty(slice_t.numel(), dtype=torch.int32, device="cuda")
    # slices start at arbitrary offsets: the kernels need a 4-byte aligned base, so copy
    buf = slice_t.clone()
    rsn._lib.check(lib.rsn_dev_lzss_match(buf.data_ptr(), buf.numel(), window, out.data_ptr(), sp))
    return out


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("variant", [0, 1])
def test_sharded_compress_matches_oracle(rsn, oracle, world, variant):
    import torch

    lib = rsn._lib.lib()
    data = synth.mixed(700000, 5, segment=100000) + synth.repetitive(150000, 9, motif=900)
    enc = oracle.escape(data)
    n = len(enc)
    stream = torch.cuda.Stream()
    sp = C.c_void_p(stream.cuda_stream)
    with torch.cuda.stream(stream):
        enc_t = torch.frombuffer(bytearray(enc), dtype=torch.uint8).cuda()
        # the escape entry point yields the same buffer
        d_in = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
        eo, en = C.c_void_p(), C.c_size_t()
        rsn._lib.check(lib.rsn_dev_lzss_escape(d_in.data_ptr(), len(data), C.byref(eo), C.byref(en), sp))
        assert en.value == n
        h = (C.c_uint8 * n)()
        rsn._lib.check(lib.rsn_dev_download(eo, n, h, sp))
        assert bytes(h) == enc
        lib.rsn_dev_free(eo, sp)

        packed = torch.zeros(n, dtype=torch.int32, device="cuda")
        for r in range(world):
            a, b, lo, hi = parallel.shard_bounds(n, world, r, 4096)
            if b > a:
                part = _gpu_match(lib, rsn, enc_t[lo:hi], 4096, sp)
                packed[a:b] = part[a - lo: b - lo]
        ln, off = oracle.lzss_match_arrays(enc, 4096, threads=8)
        got = packed.cpu().numpy().view(np.uint32)
        np.testing.assert_array_equal(got >> 16, ln)
        sel = ln >= 5  # off is only specified where a reference can be emitted (variant B: L >= 5)
        np.testing.assert_array_equal((got & 0xFFFF)[sel], off[sel])

        out, on = C.c_void_p(), C.c_size_t()
        rsn._lib.check(lib.rsn_dev_lzss_emit(enc_t.data_ptr(), n, 4096, variant, packed.data_ptr(), C.byref(out),
                                             C.byref(on), sp))
        hb = (C.c_uint8 * on.value)()
        rsn._lib.check(lib.rsn_dev_download(out, on.value, hb, sp))
        lib.rsn_dev_free(out, sp)
    want = oracle.lzss_compress_async(data, 4096, threads=8) if variant == 0 else oracle.lzss_compress_iter(data, 4096)
    assert bytes(hb) == want


@pytest.mark.parametrize("shards", [2, 3, 8])
def test_sharded_c_abi_matches_single_gpu(rsn, oracle, shards):
    """rsn_lzss_compress_sharded (BASELINE configs[4] behind the C ABI): one host thread per shard,
    shard g on device g mod the device count, so this runs on a single GPU as well.  Same bytes as the
    single-GPU call and the oracle, for data with and without escapes, ragged sizes, and inputs too
    small to give every shard a parse block."""
    cases_ = {
        "text": synth.text(300000, 31),
        "logs_with_escapes": synth.logs(250000, 32),
        "repetitive": synth.repetitive(400000, 33, motif=1500),
        "escape_heavy": (b"<\\\xff" * 30000) + synth.text(50000, 34),
        "runs": b"a" * 100000 + b"b" * 50001,
        "small": synth.text(5000, 35),
        "tiny": b"abcabcabc",
        "empty": b"",
    }
    for name, data in cases_.items():
        got = rsn.lz.CompressAsyncSharded(data, shards)
        assert got == rsn.lz.CompressAsync(data), (name, shards)
        assert got == oracle.lzss_compress_async(data, 4096, threads=8), (name, shards)
    for w in (100, 1024):
        data = cases_["logs_with_escapes"]
        got = rsn._lib.call_host(rsn._lib.lib().rsn_lzss_compress_sharded, data, w, 0, shards)
        assert got == oracle.lzss_compress_async(data, w, threads=8), (w, shards)
