#!/usr/bin/env python
"""Position-range sharded LZSS compress of ONE stream across ranks (BASELINE configs[4] shape).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29533 tools/run_sharded.py [MiB] [--check]

Every rank holds the escaped stream, computes the match arrays of its position range from a slice
with a window-sized halo and look-ahead (rsn_dev_lzss_match), the arrays are all-gathered with NCCL
(4 bytes per position over NVLink), and rank 0 runs the merge/emit (rsn_dev_lzss_emit).  Prints one
JSON line with the timings (CUDA events, max over ranks).
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import raisin_b200 as rsn  # noqa: E402
from raisin_b200 import parallel, synth  # noqa: E402


def main():
    mib = float(sys.argv[1]) if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else 256
    check = "--check" in sys.argv
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = rsn._lib.lib()
    rsn._lib.check(lib.rsn_init(local))
    n = int(mib * (1 << 20))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sp = C.c_void_p(stream.cuda_stream)
    # highly repetitive, no bytes that need escaping; large streams are tiled and mutated on the
    # device (every rank builds the same bytes) so that no rank holds several host copies
    base_n = min(n, 64 << 20)
    base = torch.frombuffer(bytearray(synth.repetitive(base_n, 5, motif=3000)), dtype=torch.uint8).cuda()
    if n > base_n:
        enc = base.repeat(-(-n // base_n))[:n].contiguous()
        idx = torch.arange(base_n, n, 65537, device="cuda")
        enc[idx] = (97 + (idx // 65537) % 26).to(torch.uint8)   # one changed byte per ~64 KiB beyond the first tile
        del idx
    else:
        enc = base
    del base
    torch.cuda.synchronize()
    W = 4096

    def match_fn(sl, window):
        out = torch.empty(sl.numel(), dtype=torch.int32, device="cuda")
        buf = sl.clone() if (sl.data_ptr() & 15) else sl
        rsn._lib.check(lib.rsn_dev_lzss_match(buf.data_ptr(), buf.numel(), window, out.data_ptr(), sp))
        return out

    times = []
    result = None
    iters = int(os.environ.get("RSN_SHARD_ITERS", "4"))
    for it in range(iters):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev[0].record(stream)
        packed = parallel.sharded_match(enc, n, W, match_fn, dist=dist if world > 1 else None, device="cuda")
        ev[1].record(stream)
        if rank == 0:
            out, on = C.c_void_p(), C.c_size_t()
            rsn._lib.check(lib.rsn_dev_lzss_emit(enc.data_ptr(), n, W, 0, packed.data_ptr(), C.byref(out), C.byref(on), sp))
            if it == iters - 1:
                hb = (C.c_uint8 * on.value)()
                rsn._lib.check(lib.rsn_dev_download(out, on.value, hb, sp))
                result = bytes(hb)
            lib.rsn_dev_free(out, sp)
        ev[2].record(stream)
        torch.cuda.synchronize()
        t = torch.tensor([ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(t.tolist())
    if rank == 0:
        m_ms, e_ms = times[-1]
        line = {"workload": "sharded single-stream lzss compress", "n_gpus": world, "stream_bytes": n,
                "match_plus_allgather_ms": m_ms, "merge_emit_ms_rank0": e_ms,
                "match_GBps": n / (m_ms * 1e-3) / 1e9, "compressed_bytes": len(result),
                "allgather_bytes_per_rank": 4 * n}
        if check:
            # the whole compress on this GPU alone, device buffers, for the comparison and the timing
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            single = None
            for it in range(2):
                out, on = C.c_void_p(), C.c_size_t()
                e0.record(stream)
                rsn._lib.check(lib.rsn_dev_lzss_compress(enc.data_ptr(), n, W, 0, C.byref(out), C.byref(on), sp))
                e1.record(stream)
                torch.cuda.synchronize()
                if it:
                    hb = (C.c_uint8 * on.value)()
                    rsn._lib.check(lib.rsn_dev_download(out, on.value, hb, sp))
                    single = bytes(hb)
                lib.rsn_dev_free(out, sp)
            line["single_gpu_compress_ms"] = e0.elapsed_time(e1)
            line["identical_to_single_gpu"] = single == result
            if n < (1 << 32):
                back = rsn.lz.Decompress(result)
                line["roundtrip_ok"] = back == bytes(enc.cpu().numpy())
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
