"""Host-side plumbing for the multi-GPU forms of the path (one process per GPU, torch.distributed).

Two shapes, both from BASELINE.json's north_star:

* independent files / streams: `partition_files` assigns whole files to ranks; no data-path
  collective (weak scaling).
* one large stream, LZSS match search sharded by position range: rank r computes the per-position
  match arrays of its range from a slice that carries a window-sized halo on the left (the search
  buffer of lzss.go:123-129) and a window-sized look-ahead on the right (a match is at most W long),
  the arrays are all-gathered (NCCL over NVLink on GPUs, gloo in the CPU tests) and the sequential
  merge of lzss.go:134-151 runs once on the gathered arrays.

The per-slice search itself is injected (`match_fn`), so the same logic is exercised on CPU with the
oracle and on GPUs with `rsn_dev_lzss_match`.
"""
from __future__ import annotations

from typing import Callable, List, Tuple


def partition_files(count: int, world: int, rank: int) -> List[int]:
    """Round-robin assignment of file indices to ranks (sizes in config 4 are equal)."""
    return list(range(rank, count, world))


def shard_bounds(n: int, world: int, rank: int, window: int, align: int = 4096) -> Tuple[int, int, int, int]:
    """(a, b, lo, hi): rank owns positions [a, b) of the escaped stream and must load [lo, hi).

    Ranges are aligned to `align` positions (the parse block) except for the last one.  The slice
    adds `window` bytes on the left (every source inside the search buffer of a position >= a starts
    at >= a - window) and `window` on the right (a match found at a position < b ends before
    b + window), both clipped to the stream.
    """
    per = -(-n // world)
    per = -(-per // align) * align
    a = min(n, rank * per)
    b = min(n, a + per)
    lo = max(0, a - window)
    hi = min(n, b + window)
    return a, b, lo, hi


def sharded_match(enc, n: int, window: int, match_fn: Callable, dist=None, device="cpu"):
    """All ranks return the packed (len << 16 | off) array of the WHOLE stream.

    `enc`: the escaped stream (a torch.uint8 tensor on `device`, identical on every rank).
    `match_fn(slice_tensor, window) -> torch.int32 tensor` of the slice's packed arrays.
    """
    import torch

    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    a, b, lo, hi = shard_bounds(n, world, rank, window)
    per = shard_bounds(n, world, 0, window)[1]  # common chunk size (the last rank may own less)
    local = torch.zeros(per, dtype=torch.int32, device=device)
    if b > a:
        packed = match_fn(enc[lo:hi], window)
        local[: b - a] = packed[a - lo: b - lo]
    if world == 1:
        return local[:n]
    gathered = torch.empty(per * world, dtype=torch.int32, device=device)
    dist.all_gather_into_tensor(gathered, local)
    return gathered[:n]
