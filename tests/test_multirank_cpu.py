"""world_size-2 tests of the multi-GPU host logic on CPU (gloo): file partitioning and the
position-range sharded match search with halo + all-gather, with the oracle standing in for the
per-slice kernel."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from raisin_b200 import parallel, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_match(slice_t, window):
    from oracle import pyoracle as po

    ln, off = po.lzss_match_arrays(slice_t.numpy().tobytes(), window)
    return torch.from_numpy(((ln.astype(np.int64) << 16) | off.astype(np.int64)).astype(np.uint32).view(np.int32))


def _worker(rank, world, port, data, window, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import pyoracle as po

        enc = po.escape(data)
        enc_t = torch.frombuffer(bytearray(enc), dtype=torch.uint8)
        packed = parallel.sharded_match(enc_t, len(enc), window, _oracle_match, dist=dist)
        ln, off = po.lzss_match_arrays(enc, window)
        want = ((ln.astype(np.int64) << 16) | off.astype(np.int64)).astype(np.uint32)
        got = packed.numpy().view(np.uint32)
        ok = bool((got == want).all())
        files = parallel.partition_files(11, world, rank)
        gathered = [None] * world
        dist.all_gather_object(gathered, files)
        flat = sorted(i for g in gathered for i in g)
        out_q.put((rank, ok, flat == list(range(11))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("window", [4096, 100])
def test_sharded_match_two_ranks(window):
    data = synth.mixed(40000, 77, segment=9000) + b"<" * 30 + synth.repetitive(9000, 3, motif=700)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, data, window, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok and part for _, ok, part in res), res


def test_shard_bounds_cover_and_halo():
    for n in (0, 1, 4095, 4096, 4097, 100000, 1 << 20):
        for world in (1, 2, 4, 8):
            covered = 0
            for r in range(world):
                a, b, lo, hi = parallel.shard_bounds(n, world, r, 4096)
                assert a == covered or (a == n and b == n)
                covered = max(covered, b)
                assert lo == max(0, a - 4096) and hi == min(n, b + 4096)
            assert covered == n


def test_partition_files():
    for world in (1, 2, 4, 8):
        seen = sorted(i for r in range(world) for i in parallel.partition_files(4096, world, r))
        assert seen == list(range(4096))
        sizes = [len(parallel.partition_files(4096, world, r)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
