"""CPU check of the necessary condition the NMS candidate kernels use to drop pairs before the clip (d3d_b200/csrc/nms.cu, nms_area_bound):
the intersection of two rotated boxes is at most ox * oy, the overlaps of box B with the bounding rectangle of box A in B's frame, so a
pair can only exceed an IoU threshold t if ox * oy >= t / (1 + t) * (area A + area B).  The arithmetic below restates the device function
in single precision, step for step; the truth is the oracle's long-double polygon clip.  The bound must never reject a pair whose true IoU
exceeds the threshold (the keep mask would change), and it should reject most pairs of near-parallel boxes that fail it (that is its use)."""
import numpy as np

from conftest import gen_boxes, proposals

F = np.float32


def area_bound_pass(A, B, thr):
    """[len(A), len(B)] bool: the device test, A = rows, B = columns (the bound is evaluated in B's frame), float32 throughout"""
    ax, ay, bx, by = F(A[:, 0])[:, None], F(A[:, 1])[:, None], F(B[:, 0])[None], F(B[:, 1])[None]
    ac, as_ = F(np.cos(A[:, 4]))[:, None], F(np.sin(A[:, 4]))[:, None]
    bc, bs = F(np.cos(B[:, 4]))[None], F(np.sin(B[:, 4]))[None]
    ahw, ahh = F(np.abs(A[:, 2]) / 2)[:, None] * F(1.000001), F(np.abs(A[:, 3]) / 2)[:, None] * F(1.000001)
    bhw, bhh = F(np.abs(B[:, 2]) / 2)[None] * F(1.000001), F(np.abs(B[:, 3]) / 2)[None] * F(1.000001)
    aarea, barea = F(A[:, 2] * A[:, 3])[:, None], F(B[:, 2] * B[:, 3])[None]
    err = F(2) * ((np.abs(ax) + np.abs(ay)) * F(2.4e-7) + (np.abs(bx) + np.abs(by)) * F(2.4e-7))
    dx, dy = ax - bx, ay - by
    px, py = bc * dx + bs * dy, bc * dy - bs * dx
    cr, sr = np.abs(ac * bc + as_ * bs), np.abs(as_ * bc - ac * bs)
    ex, ey = cr * ahw + sr * ahh + err, sr * ahw + cr * ahh + err
    ox = np.minimum(bhw, px + ex) - np.maximum(-bhw, px - ex)
    oy = np.minimum(bhh, py + ey) - np.maximum(-bhh, py - ey)
    tau = F(thr) / (F(1) + F(thr)) * F(0.999999)
    return np.maximum(ox, F(0)) * np.maximum(oy, F(0)) * F(1.001) >= tau * (aarea + barea)


def test_bound_never_drops_a_hit(oracle):
    rng = np.random.default_rng(3)
    cases = [(gen_boxes(rng, 300), gen_boxes(rng, 300)),                                   # C1 distribution: any heading, any aspect
             (proposals(rng, 400, 12, extent=20.0)[0],) * 2,                                # clustered near-parallel proposals (C3 / C5)
             (proposals(rng, 300, 8, extent=5e5)[0] + np.array([3e6, -2e6, 0, 0, 0]),) * 2]  # far from the origin: float centres are 0.25 m apart
    for A, B in cases:
        truth = oracle.iou2dr_truth(A, B)
        for thr in (0.0, 0.05, 0.3, 0.5, 0.7, 0.95):
            ok = area_bound_pass(A, B, thr)
            assert not np.any((truth > thr) & ~ok), (thr, float(truth[(truth > thr) & ~ok].max()))


def test_bound_is_tight_for_parallel_boxes(oracle):
    rng = np.random.default_rng(4)
    P = proposals(rng, 600, 20, extent=30.0)[0]
    truth = oracle.iou2dr_truth(P, P)
    near = (np.hypot(P[:, None, 0] - P[None, :, 0], P[:, None, 1] - P[None, :, 1]) < 5.0) & ~np.eye(len(P), dtype=bool)
    ok = area_bound_pass(P, P, 0.5)
    misses = near & (truth <= 0.5)
    assert misses.sum() > 1000 and (ok & misses).sum() < 0.5 * misses.sum()   # most near pairs below the threshold never reach the clip


def test_widened_circle_test_is_a_superset():
    """the single-precision circle test of the candidate kernels (centres converted to float, radii rounded up, widened by 2^-22 of the
    coordinates and 1e-5 relative) passes every pair whose bounding circles meet in double precision -- near the origin, far from it, and
    for pairs that touch to within rounding"""
    rng = np.random.default_rng(9)
    for shift in (0.0, 1e3, 1e6, 3e7):
        n = 4000
        c1 = rng.normal(0, 3, (n, 2)) + shift
        r1, r2 = rng.random(n) * 3 + 0.1, rng.random(n) * 3 + 0.1
        ang = rng.random(n) * 2 * np.pi
        gap = np.concatenate([rng.normal(0, 1e-9, n // 2), rng.normal(0, 1.0, n - n // 2)])   # half of the pairs touch to within 1e-9
        d = np.maximum(r1 + r2 + gap, 0)
        c2 = c1 + d[:, None] * np.stack([np.cos(ang), np.sin(ang)], 1)
        meet = np.hypot(c1[:, 0] - c2[:, 0], c1[:, 1] - c2[:, 1]) <= r1 + r2
        ax, ay, bx, by = F(c1[:, 0]), F(c1[:, 1]), F(c2[:, 0]), F(c2[:, 1])
        up = lambda r: np.where(F(r).astype(np.float64) >= r, F(r), np.nextafter(F(r), F(np.inf)))   # __double2float_ru
        ra = up(r1) + (np.abs(ax) + np.abs(ay)) * F(2.4e-7)
        rb = up(r2) + (np.abs(bx) + np.abs(by)) * F(2.4e-7)
        dx, dy, rs = ax - bx, ay - by, ra + rb
        passed = dx * dx + dy * dy <= rs * rs * F(1.00001)
        assert not np.any(meet & ~passed), shift
