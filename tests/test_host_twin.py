"""CPU check of the ALGORITHM the CUDA kernels run: d3d_b200/csrc/geom.cuh compiled as host code
(tests/host_twin) against the oracle.  This validates the clamped-integral clip on a box without a GPU;
the GPU tests (-m gpu) validate the kernels themselves."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import golden, gen_boxes

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def twin():
    p = os.path.join(HERE, "host_twin", "libtwin.so")
    if not os.path.exists(p):
        subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "host_twin")])
    lib = C.CDLL(p)

    def run(b1, b2, aabb=0):
        b1, b2 = np.ascontiguousarray(b1), np.ascontiguousarray(b2)
        out = np.empty((len(b1), len(b2)), b1.dtype)
        f = lib.twin_iou_f32 if b1.dtype == np.float32 else lib.twin_iou_f64
        f(b1.ctypes.data_as(C.c_void_p), C.c_long(len(b1)), b2.ctypes.data_as(C.c_void_p), C.c_long(len(b2)),
          out.ctypes.data_as(C.c_void_p), C.c_int(aabb))
        return out
    return run


def test_clip_vs_reference_and_truth(twin, oracle):
    rng = np.random.default_rng(0)
    A, B = gen_boxes(rng, 400), gen_boxes(rng, 300)
    t = twin(A, B)
    assert np.abs(t - oracle.iou2dr(A, B)).max() < 1e-10          # reference RC, generic position
    assert np.abs(t - oracle.iou2dr_truth(A, B)).max() < 1e-12
    A32, B32 = A.astype(np.float32), B.astype(np.float32)
    tr = oracle.iou2dr_truth(A32.astype(np.float64), B32.astype(np.float64))
    assert np.abs(twin(A32, B32) - tr).max() < 2e-6               # north_star fp32 tolerance is 1e-4
    assert np.array_equal(twin(A, B, 1), oracle.iou2d(A, B))      # method="box" bit-exact in fp64


def test_clip_degenerate_list(twin):
    g = golden("iou_degenerate.npz")
    for i, name in enumerate(g["names"]):
        a, b = g["boxes1"][i:i + 1], g["boxes2"][i:i + 1]
        assert abs(twin(a, b)[0, 0] - g["truth"][i]) < 1e-9, name
        assert abs(twin(a.astype(np.float32), b.astype(np.float32))[0, 0] - g["truth"][i]) < 1e-5, name


def test_clip_near_degenerate_sweep(twin, oracle):
    """same box with tiny heading / offset perturbations (and +pi flips): no blow-ups"""
    rng = np.random.default_rng(4)
    base = gen_boxes(rng, 150)
    for dr in (0, 1e-12, 1e-9, 1e-7, 1e-5):
        for dxy in (0, 1e-9, 1e-6, 1e-3):
            b2 = base.copy()
            b2[:, 4] += dr + np.pi * rng.integers(0, 2, len(base))
            b2[:, 0] += dxy
            for dt, tol in ((np.float64, 1e-9), (np.float32, 3e-5)):
                a_, b_ = base.astype(dt), b2.astype(dt)
                t = np.array([twin(a_[i:i + 1], b_[i:i + 1])[0, 0] for i in range(len(base))])
                tr = np.array([oracle.iou2dr_truth(a_[i:i + 1].astype(np.float64), b_[i:i + 1].astype(np.float64))[0, 0]
                               for i in range(len(base))])
                assert np.abs(t - tr).max() < tol, (dr, dxy, dt)
