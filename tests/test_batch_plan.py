"""CPU test of the host logic that decides which files of a batch travel together
(rsn_batch_plan = the planner rsn_batch_layers uses; no device needed)."""
import ctypes as C

import numpy as np

from raisin_b200 import _lib

MAX_FILE = 4 << 20          # kBatchMaxFile
GROUP_BYTES = 64 << 20
GROUP_FILES = 2048


def plan(sizes):
    L = _lib.lib()
    n = len(sizes)
    ns = (C.c_size_t * max(n, 1))(*sizes)
    group_of = (C.c_int64 * max(n, 1))()
    ng = C.c_size_t(0)
    assert L.rsn_batch_plan(n, ns, group_of, C.byref(ng)) == 0
    return list(group_of)[:n], ng.value


def check(sizes):
    group_of, ng = plan(sizes)
    groups = {}
    for i, (n, g) in enumerate(zip(sizes, group_of)):
        if n == 0 or n > MAX_FILE:
            assert g == -1, (i, n)
            continue
        assert 0 <= g < ng
        groups.setdefault(g, []).append(i)
    assert sorted(groups) == list(range(ng))            # no empty groups, dense numbering
    for g, idx in groups.items():
        ns = [sizes[i] for i in idx]
        assert len(idx) <= GROUP_FILES
        assert sum(ns) <= GROUP_BYTES
        assert idx == sorted(idx)                       # input order kept inside a group
        # one size class: per-file arrays are sized for the largest file of the group
        assert max(ns) < 2 * max(min(ns), 4096), (min(ns), max(ns))
        assert len(idx) * max(ns) <= 2 * max(sum(ns), len(idx) * 4096)
    return groups


def test_config4_shape_is_groups_of_256():
    groups = check([262144] * 4096)
    assert sorted(len(v) for v in groups.values()) == [256] * 16
    # a rank's share when the batch is split over 8 GPUs: one group for each of 8 workers
    groups = check([262144] * 512)
    assert sorted(len(v) for v in groups.values()) == [64] * 8


def test_mixed_sizes_do_not_share_a_group():
    sizes = [100 + 37 * (j % 50) for j in range(600)]
    sizes[17], sizes[400], sizes[599] = 3 << 20, 2 << 20, 4 << 20
    groups = check(sizes)
    big = {g for g, idx in groups.items() if any(sizes[i] >= (2 << 20) for i in idx)}
    assert len(big) == 2                                # [2 MiB, 4 MiB) together, 4 MiB alone
    assert all(sizes[i] >= (2 << 20) for g in big for i in groups[g])


def test_edges_and_random_mixes():
    check([])
    check([0, 0, 5, MAX_FILE, MAX_FILE + 1, 1, 4095, 4096, 8191, 8192])
    rng = np.random.default_rng(5)
    for _ in range(20):
        k = int(rng.integers(1, 3000))
        sizes = [int(x) for x in np.exp(rng.uniform(0, np.log(6 << 20), size=k)).astype(np.int64)]
        for j in rng.integers(0, k, size=k // 10):
            sizes[int(j)] = 0
        check(sizes)
