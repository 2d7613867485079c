"""CPU restatement of how the voxel bucket kernel finds the K smallest point indices of a crowded voxel (d3d_b200/csrc/voxel_tiles.cu,
vt_bucket_kernel: TRIM keeps the first max_points points of a voxel in index order, voxelize.cpp:288-484).  The table already holds the
smallest index; every other point runs a min-cascade over K - 1 levels -- level j keeps the smallest value it has been offered and passes
the larger one on -- and a point is kept iff it is the smallest or not above the last level.  The atomics of different points interleave
in any order: the result must not depend on it."""
import numpy as np

NONE = 0xffffffff


def kept_by_cascade(idx, K, rng):
    head = idx.min()
    rec = [NONE] * max(K - 1, 0)
    # every point's cascade is a sequence of atomicMin steps; interleave the sequences of all points at random
    state = {int(p): [0, int(p)] for p in idx if p != head}          # point -> [next level, value carried]
    live = list(state)
    while live:
        p = live[int(rng.integers(len(live)))]
        j, v = state[p]
        if j >= K - 1 or v == NONE:
            live.remove(p)
            continue
        old = rec[j]
        rec[j] = min(old, v)                                          # atomicMin
        state[p] = [j + 1, max(old, v)]                               # the larger value moves on
    return np.array([p == head or (K >= 2 and p <= rec[K - 2]) for p in idx])


def test_cascade_keeps_the_k_smallest_in_any_order():
    rng = np.random.default_rng(5)
    for trial in range(200):
        K = int(rng.integers(1, 9))
        c = int(rng.integers(K + 1, 60))                              # crowded: more than K points
        idx = rng.choice(200000, c, replace=False)
        exp = np.isin(idx, np.sort(idx)[:K])
        assert np.array_equal(kept_by_cascade(idx, K, rng), exp), (trial, K, c)
