"""d3d_b200.point -- aligned_scatter: gather per-point features from a dense feature map at fractional
coordinates, with gradient.  Mirrors reference d3d/point/__init__.py:13-67 and scatter.h:37."""
import ctypes as C
import enum

import torch

from .. import _cabi as _c


class AlignType(enum.IntEnum):
    DROP = 0
    MEAN = 1
    LINEAR = 2
    MAX = 3
    NEAREST = 4


cuda_available = True


def _dims(shape):
    d = (C.c_int64 * 3)(1, 1, 1)
    for i, s in enumerate(shape[2:]):
        d[i] = int(s)
    return d


class ScatterPlan:
    """Scratch of the tile path (2-D maps; empty where only the gather path applies).  After a forward or backward call it holds the
    binning of that call's coordinates, which a following call on the SAME coordinates and map shape reuses (the backward pass after its
    forward pass: `AlignedScatter` keeps the plan between the two)."""

    def __init__(self, n, dim, nbatch, dims, device):
        nws = int(_c.scatter_workspace_bytes(n, dim, nbatch, dims))
        self.ws = _c.workspace(nws, device) if nws else None
        self.key = (n, dim, nbatch, tuple(dims))
        self.ready = False

    def args(self, key):
        """(pointer, bytes, reuse flag) for a call on `key`"""
        if self.ws is None:
            return None, 0, 0
        if key != self.key:
            raise ValueError("scatter plan was made for another problem size")
        reuse = int(self.ready)
        self.ready = True
        return _c.ptr(self.ws), self.ws.numel(), reuse


def aligned_scatter_forward_cuda(coords, image_feature, atype, plan=None):
    """reference d3d/point/scatter.h:39-41 (plan: optional ScatterPlan shared with the backward call on the same coordinates)"""
    if int(atype) not in (AlignType.MEAN, AlignType.LINEAR):
        raise ValueError("Unsupported align type!")
    code = _c.dtype_code(image_feature.dtype)
    if coords.dtype != image_feature.dtype:
        raise RuntimeError("coordinates and feature_map must have the same dtype")
    n, dim = coords.shape[0], coords.shape[1] - 1
    if dim not in (1, 2, 3) or image_feature.dim() != dim + 2:
        raise ValueError("Unsupported dimension size: " + str(dim))
    coords, image_feature = coords.contiguous(), image_feature.contiguous()
    out = torch.empty((n, image_feature.shape[1]), dtype=image_feature.dtype, device=image_feature.device)
    dims = _dims(image_feature.shape)
    with torch.cuda.device(image_feature.device):
        if plan is None:
            plan = ScatterPlan(n, dim, image_feature.shape[0], dims, image_feature.device)
        ws, nws, reuse = plan.args((n, dim, image_feature.shape[0], tuple(dims)))
        st = _c.scatter_forward_ws(_c.ptr(coords), n, dim, _c.ptr(image_feature), image_feature.shape[0], image_feature.shape[1],
                                   dims, int(atype), code, _c.ptr(out), ws, nws, reuse, _c.stream_ptr())
    _c.check(st, "aligned_scatter_forward")
    return out


def aligned_scatter_backward_cuda(coords, grad, atype, image_grad, plan=None):
    """reference d3d/point/scatter.h:42-45: accumulates into image_grad in place"""
    code = _c.dtype_code(grad.dtype)
    n, dim = coords.shape[0], coords.shape[1] - 1
    coords, grad = coords.contiguous(), grad.contiguous()
    assert image_grad.is_contiguous()
    dims = _dims(image_grad.shape)
    with torch.cuda.device(image_grad.device):
        if plan is None:
            plan = ScatterPlan(n, dim, image_grad.shape[0], dims, image_grad.device)
        ws, nws, reuse = plan.args((n, dim, image_grad.shape[0], tuple(dims)))
        st = _c.scatter_backward_ws(_c.ptr(coords), n, dim, _c.ptr(grad), image_grad.shape[0], image_grad.shape[1],
                                    dims, int(atype), code, _c.ptr(image_grad), ws, nws, reuse, _c.stream_ptr())
    _c.check(st, "aligned_scatter_backward")


class AlignedScatter(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image_feature, coords, atype):
        coords = coords.contiguous()
        ctx.save_for_backward(coords)
        ctx.atype = atype
        ctx.image_shape = image_feature.shape
        ctx.image_dtype = image_feature.dtype
        ctx.image_device = image_feature.device
        n, dim = coords.shape[0], coords.shape[1] - 1
        ctx.plan = None
        if dim in (1, 2, 3) and image_feature.dim() == dim + 2:
            with torch.cuda.device(image_feature.device):
                ctx.plan = ScatterPlan(n, dim, image_feature.shape[0], _dims(image_feature.shape), image_feature.device)
        return aligned_scatter_forward_cuda(coords, image_feature, atype, ctx.plan)

    @staticmethod
    def backward(ctx, grad):
        coords, = ctx.saved_tensors
        image_grad = torch.zeros(ctx.image_shape, dtype=ctx.image_dtype, device=ctx.image_device)
        aligned_scatter_backward_cuda(coords, grad, ctx.atype, image_grad, ctx.plan)
        return image_grad, None, None


def aligned_scatter(coordinates, feature_map, method="drop"):
    '''
    Gather the values given coordinates in feature_map (reference d3d/point/__init__.py:41-67).

    :param feature_map: B x C x D1 x ... x Dm
    :param coordinates: N x (m+1), batch index in column 0, same dtype as feature_map
    :param method: drop | mean | linear
    :return: N x C
    '''
    method = (method or "DROP").upper()
    if method == "DROP":
        coordinates = coordinates.long()
        _, ndim = coordinates.shape
        assert len(feature_map.shape) == ndim + 1
        indexing = (coordinates[:, 0], slice(None)) + tuple(coordinates[:, i] for i in range(1, ndim))
        return feature_map[indexing]
    align_type = getattr(AlignType, method)
    odev = feature_map.device
    if odev.type != "cuda":   # host tensors: stage on the GPU, return on the host (no CPU compute path)
        out = AlignedScatter.apply(_c.to_device(feature_map), _c.to_device(coordinates), align_type)
        return out.cpu()
    return AlignedScatter.apply(feature_map, coordinates, align_type)
