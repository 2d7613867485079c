"""CPU restatement of the NMS score sort for double scores (d3d_b200/csrc/prims.cu, radix_sort_pairs_u64_hi32): a stable sort on the high
32 bits of the key, then every element takes its place inside the run of equal high halves around it by counting (low half, then
position); a run that extends more than RUN_MAX positions to either side raises the flag and the full sort runs on the sequence as it
stands.  Both branches must give the stable order of the full 64-bit key."""
import numpy as np

RUN_MAX = 16


def hi32_sort(keys):
    n = len(keys)
    hi, lo = (keys >> np.uint64(32)).astype(np.uint64), (keys & np.uint64(0xffffffff)).astype(np.uint64)
    order = np.argsort(hi, kind="stable")              # the four radix passes: stable, original index order inside equal high halves
    h, l = hi[order], lo[order]
    out = np.empty(n, np.int64)
    flag = False
    for p in range(n):                                 # rs_runfix_kernel, one thread per element
        s, e = p, p + 1
        while s > 0 and p - s < RUN_MAX and h[s - 1] == h[p]:
            s -= 1
        while e < n and e - p <= RUN_MAX and h[e] == h[p]:
            e += 1
        if (s > 0 and h[s - 1] == h[p]) or (e < n and h[e] == h[p]):
            flag = True
        at = s + sum(1 for q in range(s, e) if l[q] < l[p] or (l[q] == l[p] and q < p))
        out[at] = order[p]
    if flag:                                           # the eight passes on the sequence as it stands (a stable sort of it)
        k = keys[order]
        return order[np.argsort(k, kind="stable")], True
    return out, False


def desc_key(s):
    u = np.where(s == 0.0, 0.0, s).view(np.uint64)
    u = np.where(u >> np.uint64(63) != 0, ~u, u | np.uint64(1 << 63))
    return ~u


def test_hi32_sort_equals_full_stable_sort():
    rng = np.random.default_rng(0)
    cases = {
        "random": rng.random(3000),
        "float32 scores": rng.random(3000).astype(np.float32).astype(np.float64),
        "twins a few ulps apart": np.repeat(rng.random(1500), 2).view(np.int64) + np.tile([0, 37], 1500),
        "many equal": np.round(rng.random(3000) * 8) / 8,
        "narrow band": 0.5 + rng.random(3000) * 1e-8,
        "mixed signs and zeros": np.concatenate([rng.normal(0, 1, 1000), np.zeros(50), -np.zeros(50)]),
    }
    flags = {}
    for name, sc in cases.items():
        sc = np.asarray(sc).view(np.float64) if sc.dtype == np.int64 else np.asarray(sc, np.float64)
        keys = desc_key(sc)
        got, flag = hi32_sort(keys)
        assert np.array_equal(got, np.argsort(keys, kind="stable")), name
        flags[name] = flag
    assert not flags["random"] and not flags["twins a few ulps apart"] and flags["many equal"] and flags["narrow band"]
