"""Thin callers of the reference's OWN compiled CPU extensions (oracle/_ref/*.so).  TEST INFRASTRUCTURE ONLY.

The .so files are built from the unmodified sources under /root/reference by oracle/build_ref.py; this
module restates just enough of the reference's Python front doors (dtype policy, enum parsing, dict
plumbing -- d3d/box/__init__.py:180-276, d3d/voxel/__init__.py:16-104, d3d/point/__init__.py:13-67) to
call them with the reference's semantics.  It is used to (1) pin oracle/d3d_oracle.c, (2) generate
tests/golden/*.npz, (3) serve as the `cpu_baseline.kind == "reference"` arm of bench.py.
"""
import os

import numpy as np
import torch

from . import build_ref


def available():
    return build_ref.have_ref()


def _box():
    return build_ref.load_ref("box_impl")


def _voxel():
    return build_ref.load_ref("voxel_impl")


def _point():
    return build_ref.load_ref("point_impl")


def _t(a):
    return a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))


def box2d_iou(boxes1, boxes2, method="box", precise=True):
    b1, b2 = _t(boxes1), _t(boxes2)
    otype = b1.dtype
    if precise:
        b1, b2 = b1.to(torch.float64), b2.to(torch.float64)
    m = _box()
    if method == "box":
        r = m.iou2d_forward(b1, b2)
    elif method == "rbox":
        r = m.iou2dr_forward(b1, b2)[0]
    else:
        raise ValueError(method)
    if precise:
        r = r.to(otype)
    return r.numpy()


def crop_2dr(points, boxes):
    """the reference's own crop_2dr (d3d/box/utils.cpp:36-47): bool[M, N]"""
    return _box().crop_2dr(_t(points), _t(boxes)).numpy()


def iou_forward_backward(boxes1, boxes2, grad, method="rbox"):
    """the reference's own forward + backward of box2d_iou's four methods (d3d/box/iou.cpp), single-threaded so that its += into the box
    rows does not race: (values [N, M], grad_boxes1 [N, 5], grad_boxes2 [M, 5])"""
    m = _box()
    b1, b2, g = _t(boxes1), _t(boxes2), _t(grad)
    nt = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        if method == "box":
            v = m.iou2d_forward(b1, b2)
            g1, g2 = m.iou2d_backward(b1, b2, g)
        else:
            fw, bw = {"rbox": (m.iou2dr_forward, m.iou2dr_backward), "grbox": (m.giou2dr_forward, m.giou2dr_backward),
                      "drbox": (m.diou2dr_forward, m.diou2dr_backward)}[method]
            v, a, f = fw(b1, b2)
            g1, g2 = bw(b1, b2, g, a, f)
    finally:
        torch.set_num_threads(nt)
    return v.numpy(), g1.numpy(), g2.numpy()


_WRAP = None


def _wrap():
    """the reference's d3d/dgal_wrap.h compiled by g++ behind oracle/dgal_wrap_shim.cpp"""
    global _WRAP
    if _WRAP is None:
        import ctypes
        from . import build_ref
        if not os.path.exists(build_ref.WRAP_SO):
            build_ref.build_wrap()
        _WRAP = ctypes.CDLL(build_ref.WRAP_SO)
    return _WRAP


def box3d_iou_distance(src, dst, metric="riou"):
    """the distance cache of ScoreMatcher.prepare_boxes (d3d/tracking/matcher.pyx:45-76) computed by the reference's own
    box3dr_iou / box3d_iou (d3d/dgal_wrap.h:45-91): float32 [N, M]"""
    import ctypes
    a, b = np.array(src, dtype=np.float32, copy=True), np.array(dst, dtype=np.float32, copy=True)
    a[:, 3:6] = np.clip(a[:, 3:6], -1e3, 1e3)
    b[:, 3:6] = np.clip(b[:, 3:6], -1e3, 1e3)
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    out = np.empty((len(a), len(b)), np.float32)
    _wrap().ref_iou3d_distance(a.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(len(a)), b.ctypes.data_as(ctypes.c_void_p),
                               ctypes.c_int64(len(b)), ctypes.c_int(1 if metric == "riou" else 0), out.ctypes.data_as(ctypes.c_void_p))
    return out


def box3dr_pdist_scalar(box, points):
    """d3d/dgal_wrap.h:21-43 box3dr_pdist for one box [7] and points [N,3]: float32 [N]"""
    import ctypes
    bx, pts = np.ascontiguousarray(box, np.float32), np.ascontiguousarray(points, np.float32)
    out = np.empty(len(pts), np.float32)
    _wrap().ref_box3dr_pdist(bx.ctypes.data_as(ctypes.c_void_p), pts.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(len(pts)),
                             out.ctypes.data_as(ctypes.c_void_p))
    return out


def pdist2dr(points, boxes):
    """the reference's own pdist2dr_forward (d3d/box/dist.cpp:36-47), called with its declared argument order: (T[M, N], u8[M, N])"""
    d, ie = _box().pdist2dr_forward(_t(points), _t(boxes))
    return d.numpy(), ie.numpy()


def pdist2dr_backward(points, boxes, grad, iedge):
    """the reference's own pdist2dr_backward (d3d/box/dist.cpp:90-110): (grad_boxes [M, 5], grad_points [N, 2])"""
    gb, gp = _box().pdist2dr_backward(_t(points), _t(boxes), _t(grad), _t(iedge))
    return gb.numpy(), gp.numpy()


def box2d_nms(boxes, scores, iou_method="box", supression_method="hard", iou_threshold=0, score_threshold=0,
              supression_param=0, precise=True):
    b, s = _t(boxes), _t(scores)
    if precise:
        b, s = b.to(torch.float64), s.to(torch.float64)
    if s.dim() == 2:
        s = s.max(axis=1).values
    if b.numel() == 0:
        return np.zeros(0, bool)
    m = _box()
    sup = m.nms2d(b, s, getattr(m.IouType, iou_method.upper()), getattr(m.SupressionType, supression_method.upper()),
                  iou_threshold, score_threshold, supression_param)
    return (~sup).numpy()


class VoxelGenerator:
    def __init__(self, bounds, shape, min_points=0, max_points=30, max_voxels=20000, max_points_filter=None,
                 max_voxels_filter=None, reduction=None, dense=False):
        m = _voxel()
        self._bounds = torch.tensor(bounds, dtype=torch.float)
        self._shape = torch.tensor(shape, dtype=torch.int32)
        self._min_points, self._max_points, self._max_voxels, self._dense = min_points, max_points, max_voxels, dense
        ba = self._bounds.reshape(3, 2)
        self._size = (ba[:, 1] - ba[:, 0]) / self._shape
        dist = ba[:, 0] / self._size
        if torch.any(torch.abs(torch.round(dist) - dist) > 1e-3):
            raise ValueError("grid not aligned")
        self._offset = torch.round(dist).int()
        self._vbounds = torch.round(ba / self._size.reshape(3, 1)).long()
        self._reduction = getattr(m.ReductionType, (reduction or "NONE").upper())
        self._pf = getattr(m.MaxPointsFilterType, (max_points_filter or "NONE").upper())
        self._vf = getattr(m.MaxVoxelsFilterType, (max_voxels_filter or "NONE").upper())

    def __call__(self, points):
        m = _voxel()
        points = _t(points)
        if self._dense:
            ret = dict(m.voxelize_3d_dense(points, self._shape, self._bounds, self._max_points, self._max_voxels,
                                           self._reduction))
        else:
            sp = m.voxelize_3d_sparse(points, self._size, 3)
            ret = dict(m.voxelize_3d_filter(points, sp["points_mapping"], sp["coords"], sp["voxel_npoints"],
                                            self._vbounds, self._min_points, self._max_points, self._max_voxels,
                                            self._pf, self._vf))
            ret["coords"] = ret["coords"] - self._offset
        return {k: v.numpy() for k, v in ret.items()}


def aligned_scatter_forward(coords, feature_map, method):
    m = _point()
    return m.aligned_scatter_forward(_t(coords), _t(feature_map), getattr(m.AlignType, method.upper())).numpy()


def aligned_scatter_backward(coords, grad, method, image_shape):
    m = _point()
    g = _t(grad)
    ig = torch.zeros(tuple(image_shape), dtype=g.dtype)
    m.aligned_scatter_backward(_t(coords), g, getattr(m.AlignType, method.upper()), ig)
    return ig.numpy()
