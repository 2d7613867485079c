"""CPU restatement of the two barrier-free NMS resolves (d3d_b200/csrc/nms.cu): greedy NMS keeps a box iff no KEPT box of higher score
overlaps it.  nms_pull_kernel lets every box poll its higher-scored overlapping boxes; the per-frame kernel of the batched NMS pushes:
a kept box marks its successors suppressed, a suppressed box takes itself off their open-predecessor counts, a box whose count reaches
zero while it is still undecided is kept.  Both must reach the greedy mask whatever the order in which boxes get their turn."""
import numpy as np

UND, KEPT, SUP = 0, 1, 2


def greedy(n, edges, valid):
    keep = np.zeros(n, bool)
    sup = ~valid.copy()
    succ = [[] for _ in range(n)]
    for a, b in edges:
        succ[a].append(b)
    for i in range(n):                      # score order
        if not sup[i]:
            keep[i] = True
            for j in succ[i]:
                sup[j] = True
    return keep


def pulled(n, edges, valid, rng):
    pred = [[] for _ in range(n)]
    for a, b in edges:
        if valid[a]:                        # boxes at or below the score threshold never suppress anything: left out
            pred[b].append(a)
    st = np.where(valid, UND, SUP)
    while (st == UND).any():
        for j in rng.permutation(n):        # any schedule
            if st[j] != UND:
                continue
            ps = st[pred[j]] if pred[j] else np.zeros(0, int)
            if (ps == KEPT).any():
                st[j] = SUP
            elif not (ps == UND).any():
                st[j] = KEPT
    return st == KEPT


def pushed(n, edges, valid, rng):
    succ = [[] for _ in range(n)]
    pend = np.zeros(n, int)
    for a, b in edges:
        if valid[a]:
            succ[a].append(b)
            pend[b] += 1
    st = np.where(valid, UND, SUP)
    todo = set(np.flatnonzero(valid))       # boxes that have not reported to their successors yet
    while todo:
        for j in rng.permutation(sorted(todo)):
            if st[j] == UND and pend[j] == 0:
                st[j] = KEPT
            if st[j] != UND:
                for d in succ[j]:
                    if st[j] == KEPT:
                        st[d] = SUP         # d still counts this box: it cannot have been kept
                    else:
                        pend[d] -= 1
                todo.discard(j)
    return st == KEPT


def test_resolves_reach_the_greedy_mask():
    rng = np.random.default_rng(1)
    for trial in range(30):
        n = int(rng.integers(1, 200))
        m = int(rng.integers(0, 6 * n))
        a, b = rng.integers(0, n, m), rng.integers(0, n, m)
        edges = sorted({(min(x, y), max(x, y)) for x, y in zip(a, b) if x != y})   # higher score -> lower score
        if trial % 5 == 0:
            edges = sorted(set(edges) | {(i, i + 1) for i in range(n - 1)})          # a chain of alternating decisions as deep as the frame
        valid = rng.random(n) > 0.1
        exp = greedy(n, edges, valid)
        assert np.array_equal(pulled(n, edges, valid, rng), exp), trial
        assert np.array_equal(pushed(n, edges, valid, rng), exp), trial
