"""CPU restatement of the sorting network nmsb_morton_kernel runs on a frame's 1024..4096 Morton keys (d3d_b200/csrc/nms.cu): element
e = r * 1024 + tid lives in register r of thread tid; a compare-exchange step (k, j) pairs e with e ^ j and keeps the minimum at the
element whose bit j is clear iff bit k of e is clear.  Steps with j >= 1024 pair two registers of one thread -- the kernel writes them out
by hand: (v0, v1) ascending and (v2, v3) ascending unless k = 2048 for j = 1024; (v0, v2), (v1, v3) ascending for j = 2048 -- steps with
j < 32 run as shuffles and the rest through shared memory, all with the generic rule.  The frame order only decides which tiles the
batched NMS can skip, never a keep mask, so the GPU parity tests cannot see a wrong direction flag: this test can."""
import numpy as np


def network(keys):
    npow = len(keys)
    R = npow // 1024
    v = keys.reshape(R, 1024).copy()          # v[r][tid]
    tid = np.arange(1024)

    def cx(a, b, up):                          # the kernel's lambda on two registers of every thread
        lo, hi = np.minimum(v[a], v[b]), np.maximum(v[a], v[b])
        v[a], v[b] = (lo, hi) if up else (hi, lo)

    k = 2
    while k <= npow:
        j = k >> 1
        while j > 0:
            if j >= 1024:
                if j == 1024:
                    cx(0, 1, True)
                    if R > 2:
                        cx(2, 3, k != 2048)
                else:
                    cx(0, 2, True); cx(1, 3, True)
            else:                              # shuffle (j < 32) and shared-memory steps: the generic rule
                new = v.copy()
                for r in range(R):
                    e = r * 1024 + tid
                    o = v[r][tid ^ j]          # partner e ^ j: same register, lane / thread tid ^ j
                    takemin = ((e & j) == 0) == ((e & k) == 0)
                    new[r] = np.where(takemin, np.minimum(v[r], o), np.maximum(v[r], o))
                v = new
            j >>= 1
        k <<= 1
    return v.reshape(-1)


def test_network_sorts():
    rng = np.random.default_rng(2)
    for npow in (1024, 2048, 4096):
        for trial in range(3):
            n = int(rng.integers(npow // 2 + 1, npow + 1))
            cell = rng.integers(0, 1 << 20, n).astype(np.uint64)
            keys = np.full(npow, 0xffffffff, np.uint64)
            keys[:n] = (cell << np.uint64(12)) | np.arange(n, dtype=np.uint64)   # unique: cell << 12 | index
            out = network(keys)
            assert np.array_equal(out, np.sort(keys)), (npow, trial)
