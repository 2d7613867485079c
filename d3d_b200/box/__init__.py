"""d3d_b200.box -- pairwise IoU and NMS on (rotated) 2-D boxes.  Mirrors reference d3d/box/__init__.py
(box2d_iou :180-224, box2d_nms :226-276, enums d3d/box/common.h:5-10) on top of the C ABI."""
import enum

import numpy as np
import torch

from .. import _cabi as _c


class IouType(enum.IntEnum):   # d3d/box/common.h:5-9
    NA = 0
    BOX = 1
    RBOX = 2
    GBOX = 3
    GRBOX = 4
    DBOX = 5
    DRBOX = 6


class SupressionType(enum.IntEnum):   # d3d/box/common.h:10 (reference spelling)
    HARD = 0
    LINEAR = 1
    GAUSSIAN = 2


cuda_available = True


def iou2d_forward_cuda(boxes1, boxes2):
    """AABB IoU of rotated boxes; CUDA tensors [N,5],[M,5] -> [N,M] (reference d3d/box/iou.h:7-9)."""
    return _pairwise(boxes1, boxes2, _c.iou2d)


def iou2dr_forward_cuda(boxes1, boxes2):
    """Rotated IoU; CUDA tensors [N,5],[M,5] -> [N,M] (reference d3d/box/iou.h:14-16; the backward-only
    nx/xflags outputs are not produced)."""
    return _pairwise(boxes1, boxes2, _c.iou2dr)


def giou2dr_forward_cuda(boxes1, boxes2):
    """Rotated GIoU; CUDA tensors [N,5],[M,5] -> [N,M] (reference d3d/box/iou.h:24-26)"""
    return _pairwise_ex(boxes1, boxes2, _c.giou2dr)


def diou2dr_forward_cuda(boxes1, boxes2):
    """Rotated DIoU; CUDA tensors [N,5],[M,5] -> [N,M] (reference d3d/box/iou.h:34-36)"""
    return _pairwise_ex(boxes1, boxes2, _c.diou2dr)


def iou2d_backward_cuda(boxes1, boxes2, grad):
    """(grad_boxes1, grad_boxes2) of the AABB IoU (reference d3d/box/iou.h:10-12)"""
    return _pairwise_backward("iou2d", boxes1, boxes2, grad)


def iou2dr_backward_cuda(boxes1, boxes2, grad, nx=None, xflags=None):
    """(grad_boxes1, grad_boxes2) of the rotated IoU (reference d3d/box/iou.h:17-20); nx / xflags are accepted and ignored"""
    return _pairwise_backward("iou2dr", boxes1, boxes2, grad)


def giou2dr_backward_cuda(boxes1, boxes2, grad, nxm=None, flags=None):
    """(grad_boxes1, grad_boxes2) of the rotated GIoU (reference d3d/box/iou.h:27-30)"""
    return _pairwise_backward("giou2dr", boxes1, boxes2, grad)


def diou2dr_backward_cuda(boxes1, boxes2, grad, nxd=None, flags=None):
    """(grad_boxes1, grad_boxes2) of the rotated DIoU (reference d3d/box/iou.h:37-40)"""
    return _pairwise_backward("diou2dr", boxes1, boxes2, grad)


def _pairwise(b1, b2, table, out=None):
    code = _c.dtype_code(b1.dtype)
    if b2.dtype != b1.dtype:
        raise RuntimeError("boxes1 and boxes2 must have the same dtype")
    n, m = b1.shape[0], b2.shape[0]
    if out is None:
        out = torch.empty((n, m), dtype=b1.dtype, device=b1.device)
    if n == 0 or m == 0:
        return out
    ws = _c.workspace(_c.iou_workspace_bytes(n, m, code), b1.device)
    with torch.cuda.device(b1.device):
        st = table[code](_c.ptr(b1), n, _c.ptr(b2), m, _c.ptr(out), out.stride(0), _c.ptr(ws), ws.numel(), _c.stream_ptr())
    _c.check(st, "box2d_iou")
    return out


def _pairwise_ex(b1, b2, table):
    """GIoU / DIoU forward: no workspace"""
    code = _c.dtype_code(b1.dtype)
    if b2.dtype != b1.dtype:
        raise RuntimeError("boxes1 and boxes2 must have the same dtype")
    n, m = b1.shape[0], b2.shape[0]
    out = torch.empty((n, m), dtype=b1.dtype, device=b1.device)
    if n and m:
        with torch.cuda.device(b1.device):
            st = table[code](_c.ptr(b1), n, _c.ptr(b2), m, _c.ptr(out), out.stride(0), _c.stream_ptr())
        _c.check(st, "box2d_iou")
    return out


def _pairwise_backward(name, b1, b2, grad):
    code = _c.dtype_code(b1.dtype)
    n, m = b1.shape[0], b2.shape[0]
    grad = grad.contiguous()
    g1, g2 = torch.empty_like(b1), torch.empty_like(b2)
    with torch.cuda.device(b1.device):
        st = _c.iou_backward[name][code](_c.ptr(b1), n, _c.ptr(b2), m, _c.ptr(grad), max(m, 1), _c.ptr(g1), _c.ptr(g2), _c.stream_ptr())
    _c.check(st, "box2d_iou backward")
    return g1, g2


def _make_iou_function(name, doc, forward_impl):
    class _F(torch.autograd.Function):
        @staticmethod
        def forward(ctx, boxes1, boxes2):
            boxes1, boxes2 = boxes1.contiguous(), boxes2.contiguous()
            ctx.save_for_backward(boxes1, boxes2)   # the reference also saves nx / xflags (9 B per pair); the backward here recomputes
            return forward_impl(boxes1, boxes2)

        @staticmethod
        def backward(ctx, grad):
            boxes1, boxes2 = ctx.saved_tensors
            return _pairwise_backward(name, boxes1, boxes2, grad)
    _F.__doc__ = doc
    return _F


Iou2D = _make_iou_function("iou2d", "Differentiable IoU function for 2D axis-aligned boxes (reference d3d/box/__init__.py:38-61)",
                           lambda a, b: _pairwise(a, b, _c.iou2d))
Iou2D.__name__ = Iou2D.__qualname__ = "Iou2D"
Iou2DR = _make_iou_function("iou2dr", "Differentiable rotated IoU function for 2D boxes (reference d3d/box/__init__.py:63-84)",
                            lambda a, b: _pairwise(a, b, _c.iou2dr))
Iou2DR.__name__ = Iou2DR.__qualname__ = "Iou2DR"
GIou2DR = _make_iou_function("giou2dr", "Differentiable rotated GIoU function for 2D boxes (reference d3d/box/__init__.py:86-107)",
                             lambda a, b: _pairwise_ex(a, b, _c.giou2dr))
GIou2DR.__name__ = GIou2DR.__qualname__ = "GIou2DR"
DIou2DR = _make_iou_function("diou2dr", "Differentiable rotated DIoU function for 2D boxes (reference d3d/box/__init__.py:109-130)",
                             lambda a, b: _pairwise_ex(a, b, _c.diou2dr))
DIou2DR.__name__ = DIou2DR.__qualname__ = "DIou2DR"


def box2d_iou(boxes1, boxes2, method="box", precise=True):
    '''
    Differentiable IoU on axis-aligned or rotated 2D boxes (reference d3d/box/__init__.py:180-224)

    :param boxes1: Input boxes, shape is N x 5 (x,y,w,h,r)
    :param boxes2: Input boxes, shape is M x 5 (x,y,w,h,r)
    :param method: 'box' - axis-aligned box, 'rbox' - rotated box,
        'grbox' - giou for rotated box, 'drbox' - diou for rotated box
    :param precise: force using double precision to calculate iou
    '''
    convert_numpy = False
    if isinstance(boxes1, np.ndarray):
        assert isinstance(boxes2, np.ndarray), "Input should be both numpy tensor or pytorch tensor!"
        boxes1 = torch.from_numpy(boxes1)
        boxes2 = torch.from_numpy(boxes2)
        convert_numpy = True

    otype = boxes1.dtype
    odev = boxes1.device
    if len(boxes1.shape) != 2 or len(boxes2.shape) != 2:
        raise ValueError("Input of rbox_2d_iou should be Nx2 tensors!")
    if boxes1.shape[1] != 5 or boxes2.shape[1] != 5:
        raise ValueError("Input boxes should have 5 fields: x, y, w, h, r")

    iou_type = getattr(IouType, method.upper())
    if iou_type == IouType.BOX:
        impl = Iou2D
    elif iou_type == IouType.RBOX:
        impl = Iou2DR
    elif iou_type == IouType.GRBOX:
        impl = GIou2DR
    elif iou_type == IouType.DRBOX:
        impl = DIou2DR
    else:
        raise ValueError("Unrecognized iou type!")

    b1, b2 = _c.to_device(boxes1), _c.to_device(boxes2)
    if precise:
        b1, b2 = b1.to(torch.float64), b2.to(torch.float64)
    result = impl.apply(b1, b2)
    if precise:
        result = result.to(otype)
    if not odev.type == "cuda":
        result = result.cpu()
    if convert_numpy:
        return result.detach().numpy()
    return result


def box2d_nms_batch(boxes, scores, offsets=None, iou_method="box", iou_threshold=0, score_threshold=0, precise=True, offsets_dev=None):
    '''
    Hard NMS on a batch of frames in one launch sequence (the reference has no batch form; per frame the result equals
    :func:`box2d_nms`).  Two call forms: lists of per-frame ``boxes`` / ``scores`` (returns a list of keep masks), or the frames
    packed back to back as ``boxes`` [total, 5] / ``scores`` [total] with ``offsets`` (int64 [nframes + 1], CPU tensor) (returns one
    keep mask bool[total]).  Frames of more than 8192 boxes fall back to one :func:`box2d_nms` call per frame.

    :param iou_method: 'box' - axis-aligned box, 'rbox' - rotated box
    :param offsets_dev: the same offsets already on the device (saves the copy)
    '''
    as_list = isinstance(boxes, (list, tuple))
    if as_list:
        if len(boxes) != len(scores):
            raise ValueError("Numbers of boxes and scores are inconsistent!")
        if len(boxes) == 0:
            return []
        tt = lambda x: torch.from_numpy(x) if isinstance(x, np.ndarray) else x
        numpy_in = isinstance(boxes[0], np.ndarray)
        blist, slist = [tt(b) for b in boxes], [tt(s_) for s_ in scores]
        odev = blist[0].device
        lens = [int(b.shape[0]) for b in blist]
        offsets = torch.zeros(len(lens) + 1, dtype=torch.int64)
        offsets[1:] = torch.tensor(lens).cumsum(0)
        boxes = torch.cat([_c.to_device(b).reshape(-1, 5) for b in blist], 0)
        scores = torch.cat([_c.to_device(s_).reshape(-1) for s_ in slist], 0)
    else:
        numpy_in = isinstance(boxes, np.ndarray)
        if numpy_in:
            boxes, scores = torch.from_numpy(boxes), torch.from_numpy(scores)
        odev = boxes.device
        if offsets is None:
            raise ValueError("packed boxes need the frame offsets")
        offsets = torch.as_tensor(offsets).to(torch.int64).cpu()
    if len(boxes) != len(scores) or int(offsets[-1]) != len(boxes):
        raise ValueError("Numbers of boxes and scores are inconsistent!")
    iou_type = getattr(IouType, iou_method.upper())
    if iou_type not in (IouType.BOX, IouType.RBOX):
        raise ValueError("Unsupported iou type!")
    b, s_ = _c.to_device(boxes), _c.to_device(scores)
    if precise:
        b, s_ = b.to(torch.float64), s_.to(torch.float64)
    elif s_.dtype != b.dtype:
        s_ = s_.to(b.dtype)
    b, s_ = b.contiguous(), s_.contiguous()
    total, nframes = int(b.shape[0]), int(offsets.numel()) - 1
    max_frame = int((offsets[1:] - offsets[:-1]).max()) if nframes > 0 else 0
    suppressed = torch.zeros(total, dtype=torch.uint8, device=b.device)
    if total > 0 and max_frame > 8192:
        for f in range(nframes):
            lo, hi = int(offsets[f]), int(offsets[f + 1])
            if hi > lo:
                suppressed[lo:hi] = nms2d_cuda(b[lo:hi], s_[lo:hi], iou_type, SupressionType.HARD, iou_threshold, score_threshold, 0).view(torch.uint8)
    elif total > 0:
        code = _c.dtype_code(b.dtype)
        od = offsets.to(b.device, non_blocking=True) if offsets_dev is None else offsets_dev
        ws = _c.workspace(_c.nms_batch_workspace_bytes(total, nframes, max_frame, code), b.device)
        with torch.cuda.device(b.device):
            st = _c.nms2d_batch[code](_c.ptr(b), _c.ptr(s_), total, _c.ptr(od), nframes, max_frame, int(iou_type), int(SupressionType.HARD),
                                      float(iou_threshold), float(score_threshold), _c.ptr(suppressed), _c.ptr(ws), ws.numel(), _c.stream_ptr())
        _c.check(st, "box2d_nms_batch")
    keep = ~suppressed.view(torch.bool)
    if odev.type != "cuda":
        keep = keep.cpu()
    if as_list:
        out = [keep[int(offsets[f]):int(offsets[f + 1])] for f in range(nframes)]
        return [k.cpu().numpy() for k in out] if numpy_in else out
    return keep.cpu().numpy() if numpy_in else keep


def crop_2dr_cuda(points, boxes):
    """bool[M, N] mask of the points [N,2] inside the rotated boxes [M,5] (reference crop_2dr, d3d/box/utils.h:45; CUDA tensors)"""
    code = _c.dtype_code(points.dtype)
    if boxes.dtype != points.dtype:
        raise RuntimeError("points and boxes must have the same dtype")
    n, m = points.shape[0], boxes.shape[0]
    mask = torch.empty((m, n), dtype=torch.bool, device=points.device)
    if n and m:
        ws = _c.workspace(_c.crop_workspace_bytes(n, m, code), points.device)
        with torch.cuda.device(points.device):
            st = _c.crop2dr[code](_c.ptr(points), n, _c.ptr(boxes), m, _c.ptr(mask), _c.ptr(ws), ws.numel(), _c.stream_ptr())
        _c.check(st, "box2dr_crop")
    return mask


def box2dr_crop(points, boxes):
    '''
    Crop point points points out given rotated boxes (reference d3d/box/__init__.py:278-287).
    The result is the M x N indicator mask of the points lying in each box (what the reference's crop_2dr returns).

    :param points: The input point points, shape: N x 2
    :param boxes: Input boxes array, shape: M x 5
    '''
    if len(points.shape) != 2 or points.shape[1] != 2:
        raise ValueError("Input points should be Nx2 tensors!")
    if len(boxes.shape) != 2 or boxes.shape[1] != 5:
        raise ValueError("Input boxes should have 5 fields: x, y, w, h, r")
    odev = points.device
    r = crop_2dr_cuda(_c.to_device(points).contiguous(), _c.to_device(boxes).contiguous())
    return r if odev.type == "cuda" else r.cpu()


def box3dp_crop(points, boxes, project_axis=2):
    '''
    Crop point points points out given rotated boxes with boxes projected to given axis (reference d3d/box/__init__.py:289-314)

    :param points: The input point points, shape: N x 3
    :param boxes: Input boxes array, shape: M x 7
    :param project_axis: Axis for the box to be projected to. {0: x, 1: y, 2: z}
    '''
    if project_axis == 0:
        points_2d, boxes_2d = points[:, [1, 2]], boxes[:, [1, 2, 4, 5, 6]]
    elif project_axis == 1:
        points_2d, boxes_2d = points[:, [0, 2]], boxes[:, [0, 2, 3, 5, 6]]
    elif project_axis == 2:
        points_2d, boxes_2d = points[:, [0, 1]], boxes[:, [0, 1, 3, 4, 6]]
    else:
        raise ValueError("The projection axis can only be 0-x, 1-y and 2-z!")
    mask_2d = box2dr_crop(points_2d, boxes_2d)
    points_p = points[:, [project_axis]].t()
    boxes_p = boxes[:, [project_axis]]
    boxes_pd = boxes[:, [3 + project_axis]] / 2
    mask_p = (points_p - boxes_pd < boxes_p) & (boxes_p < points_p + boxes_pd)
    return mask_2d & mask_p.to(mask_2d.device)


class PDist2DR(torch.autograd.Function):
    """Differentiable signed distance from points to rotated boxes (reference d3d/box/__init__.py:149-166 over pdist2dr_forward /
    pdist2dr_backward[_cuda], d3d/box/dist.h:7-23).  Returns T[M boxes, N points], positive inside -- the layout of the native function
    (and of box2dr_crop).  The reference's wrapper hands (boxes, points) to a native function declared (points, boxes) and cannot run."""
    @staticmethod
    def forward(ctx, points, boxes):
        code = _c.dtype_code(points.dtype)
        if boxes.dtype != points.dtype:
            raise RuntimeError("points and boxes must have the same dtype")
        n, m = points.shape[0], boxes.shape[0]
        dist = torch.empty((m, n), dtype=points.dtype, device=points.device)
        if n and m:
            with torch.cuda.device(points.device):
                st = _c.pdist2dr[code](_c.ptr(points), n, _c.ptr(boxes), m, _c.ptr(dist), None, _c.stream_ptr())
            _c.check(st, "box2dr_pdist")
        ctx.save_for_backward(points, boxes)
        return dist

    @staticmethod
    def backward(ctx, grad):
        points, boxes = ctx.saved_tensors
        code = _c.dtype_code(points.dtype)
        n, m = points.shape[0], boxes.shape[0]
        grad = grad.contiguous()
        grad_boxes, grad_points = torch.zeros_like(boxes), torch.zeros_like(points)
        if n and m:
            with torch.cuda.device(points.device):
                st = _c.pdist2dr_backward[code](_c.ptr(points), n, _c.ptr(boxes), m, _c.ptr(grad), _c.ptr(grad_boxes), _c.ptr(grad_points), _c.stream_ptr())
            _c.check(st, "box2dr_pdist backward")
        return grad_points, grad_boxes


def seg1d_pdist(points, segs):
    """Signed distance from points [N] to 1-D segments [M, 2] (centre, width), positive inside: [M, N]
    (reference d3d/box/__init__.py:316-328, in the [segments, points] layout of box2dr_pdist)."""
    assert torch.all(segs[:, 1] > 0)
    dsegs = (segs[:, 1] / 2)[:, None]
    ctr = segs[:, 0][:, None]
    return torch.where(points[None, :] > ctr, (ctr + dsegs) - points[None, :], points[None, :] - (ctr - dsegs))


def box2dr_pdist(points, boxes, method="rbox"):
    '''
    Calculate signed distance from points to 2d boxes (surfaces), positive inside (reference d3d/box/__init__.py:330-346).
    Differentiable in points and boxes.

    :param points: target points, shape: N x 2
    :param boxes: target boxes, shape: M x 5
    :param method: 'rbox' - rotated box
    :return: M x N
    '''
    if len(boxes.shape) != 2:
        raise ValueError("Input boxes should be Nx2 tensors!")
    if boxes.shape[1] != 5:
        raise ValueError("Input boxes should have 5 fields: x, y, w, h, r")
    if len(points.shape) != 2 or points.shape[1] != 2:
        raise ValueError("Input points should be Nx2 tensors!")
    if method != "rbox":
        raise ValueError("Only supported rotated boxes by now!")
    odev = points.device
    r = PDist2DR.apply(_c.to_device(points).contiguous(), _c.to_device(boxes).contiguous())
    return r if odev.type == "cuda" else r.cpu()


def box3dr_pdist(points, boxes, project_axis=2):
    '''
    Calculate signed distance from points to 3d boxes (surfaces) (reference d3d/box/__init__.py:348-381)

    :param points: target points, shape: N x 3
    :param boxes: target boxes, shape: M x 7
    :param project_axis: Axis for the box to be projected to. {0: x, 1: y, 2: z}
    :return: M x N
    '''
    if project_axis == 0:
        points_2d, boxes_2d = points[:, [1, 2]], boxes[:, [1, 2, 4, 5, 6]]
    elif project_axis == 1:
        points_2d, boxes_2d = points[:, [0, 2]], boxes[:, [0, 2, 3, 5, 6]]
    elif project_axis == 2:
        points_2d, boxes_2d = points[:, [0, 1]], boxes[:, [0, 1, 3, 4, 6]]
    else:
        raise ValueError("The projection axis can only be 0-x, 1-y and 2-z!")
    dist_2d = box2dr_pdist(points_2d, boxes_2d)
    dist_p = seg1d_pdist(points[:, project_axis], boxes[:, [project_axis, 3 + project_axis]]).to(dist_2d.device)
    return torch.where(dist_p > 0,
                       torch.where(dist_2d > 0, torch.min(dist_p, dist_2d), dist_2d),
                       torch.where(dist_2d > 0, dist_p, -torch.sqrt(dist_2d.square() + dist_p.square())))


def box3d_iou_distance(src_boxes, dst_boxes, metric="riou"):
    '''
    Distance matrix of the detection evaluator / tracking matcher: ``1 - iou2d * ziou`` in float32, the array
    ``ScoreMatcher.prepare_boxes`` caches (reference d3d/tracking/matcher.pyx:24-82 over box3dr_iou / box3d_iou,
    d3d/dgal_wrap.h:45-91).

    :param src_boxes: boxes to match, shape N x 7: x, y, z, lx, ly, lz, rz (the columns 2..8 of ``Target3DArray.to_numpy()``)
    :param dst_boxes: fixed boxes (such as ground truth boxes), shape M x 7
    :param metric: 'riou' - rotated BEV IoU times the z overlap ratio (DistanceTypes.RIoU), 'iou' - IoU of the BEV
        axis-aligned bounding boxes times the z overlap ratio (DistanceTypes.IoU)
    :return: float32 N x M; numpy in -> numpy out, host tensors in -> host tensor out
    '''
    convert_numpy = False
    if isinstance(src_boxes, np.ndarray):
        assert isinstance(dst_boxes, np.ndarray), "Input should be both numpy tensor or pytorch tensor!"
        src_boxes, dst_boxes = torch.from_numpy(src_boxes), torch.from_numpy(dst_boxes)
        convert_numpy = True
    if len(src_boxes.shape) != 2 or len(dst_boxes.shape) != 2 or src_boxes.shape[1] != 7 or dst_boxes.shape[1] != 7:
        raise ValueError("Input boxes should be Nx7 tensors: x, y, z, lx, ly, lz, rz")
    if metric.lower() not in ("riou", "iou"):
        raise ValueError("Unrecognized distance metric!")
    odev = src_boxes.device
    b1 = _c.to_device(src_boxes).to(torch.float32).contiguous().clone()
    b2 = _c.to_device(dst_boxes).to(torch.float32).contiguous().clone()
    b1[:, 3:6].clamp_(-1e3, 1e3)   # "prevent really weird boxes with unusual size" (matcher.pyx:50-52)
    b2[:, 3:6].clamp_(-1e3, 1e3)
    n, m = b1.shape[0], b2.shape[0]
    out = torch.empty((n, m), dtype=torch.float32, device=b1.device)
    if n and m:
        ws = _c.workspace(_c.iou3d_distance_workspace_bytes(n, m), b1.device)
        with torch.cuda.device(b1.device):
            st = _c.iou3d_distance(_c.ptr(b1), n, _c.ptr(b2), m, 1 if metric.lower() == "riou" else 0, _c.ptr(out), out.stride(0),
                                   _c.ptr(ws), ws.numel(), _c.stream_ptr())
        _c.check(st, "box3d_iou_distance")
    if odev.type != "cuda":
        out = out.cpu()
    return out.numpy() if convert_numpy else out


def nms2d_cuda(boxes, scores, iou_type, supression_type, iou_threshold, score_threshold, supression_param):
    """Suppressed mask bool[N] in original order (reference d3d/box/nms.h:6-10, nms_cuda.cu:217-244)."""
    code = _c.dtype_code(boxes.dtype)
    if scores.dtype != boxes.dtype:
        raise RuntimeError("boxes and scores must have the same dtype")
    n = boxes.shape[0]
    suppressed = torch.empty(n, dtype=torch.uint8, device=boxes.device)
    ws = _c.workspace(_c.nms_workspace_bytes(n, code), boxes.device)
    with torch.cuda.device(boxes.device):
        st = _c.nms2d[code](_c.ptr(boxes), _c.ptr(scores), n, int(iou_type), int(supression_type), float(iou_threshold),
                            float(score_threshold), float(supression_param), _c.ptr(suppressed), _c.ptr(ws), ws.numel(),
                            _c.stream_ptr())
    if st == _c.ERR_INVALID and int(iou_type) not in (IouType.BOX, IouType.RBOX):
        raise ValueError("Unsupported iou type!")
    _c.check(st, "box2d_nms")
    return suppressed.view(torch.bool)


def box2d_nms(boxes, scores, iou_method="box", supression_method="hard",
    iou_threshold=0, score_threshold=0, supression_param=0, precise=True):
    '''
    NMS on axis-aligned or rotated 2D boxes (reference d3d/box/__init__.py:226-276).  Returns the
    KEEP mask bool[N] in the original box order.

    :param iou_method: 'box' - axis-aligned box, 'rbox' - rotated box
    :param precise: force using double precision to calculate iou
    :param iou_threshold: IoU threshold for two boxes to be considered as overlapped
    :param score_threshold: Minimum score for a box to be considered as valid
    '''
    convert_numpy = False
    if isinstance(boxes, np.ndarray):
        assert isinstance(scores, np.ndarray), "Input should be both numpy tensor or pytorch tensor!"
        boxes = torch.from_numpy(boxes)
        scores = torch.from_numpy(scores)
        convert_numpy = True
    odev = boxes.device

    if len(boxes) != len(scores):
        raise ValueError("Numbers of boxes and scores are inconsistent!")
    if boxes.numel() == 0:
        mask = torch.tensor([], dtype=torch.bool)
        return mask.numpy() if convert_numpy else mask

    iou_type = getattr(IouType, iou_method.upper())
    supression_type = getattr(SupressionType, supression_method.upper())

    b, s = _c.to_device(boxes), _c.to_device(scores)
    if precise:
        b, s = b.to(torch.float64), s.to(torch.float64)
    elif s.dtype != b.dtype:
        s = s.to(b.dtype)
    if len(s.shape) == 2:
        s = s.max(axis=1).values.contiguous()

    suppressed = nms2d_cuda(b, s, iou_type, supression_type, iou_threshold, score_threshold, supression_param)
    mask = ~suppressed
    if odev.type != "cuda":
        mask = mask.cpu()
    if convert_numpy:
        return mask.numpy()
    return mask


def match_greedy(distance, src_scores, src_tags, dst_tags, thresholds):
    '''
    Greedy score-ordered matching on a distance matrix: ``ScoreMatcher.match`` of the reference (d3d/tracking/matcher.pyx:138-162 over
    match_by_order :93-122) for one or many threshold sets at once -- the detection evaluator's loop over score thresholds
    (d3d/benchmarks.pyx:220-238) is one launch.  Source boxes are visited from the best score down (ties: the later box first, like
    ``np.flip(np.argsort(scores))``); each takes the closest free destination box of its category whose distance is <= the threshold.

    :param distance: float32 N x M, e.g. the result of :func:`box3d_iou_distance`
    :param src_scores: N scores of the source boxes
    :param src_tags, dst_tags: integer category per source / destination box, values in [0, C)
    :param thresholds: T x C (or C for a single set): maximum distance per category
    :return: (src_assignment int32 T x N, dst_assignment int32 T x M) with -1 for unmatched boxes; the leading dimension is dropped
        for a single threshold set.  numpy in -> numpy out.
    '''
    convert_numpy = isinstance(distance, np.ndarray)
    as_t = lambda x, dt: (torch.from_numpy(np.ascontiguousarray(x)) if isinstance(x, np.ndarray) else torch.as_tensor(x)).to(dt)
    distance = as_t(distance, torch.float32)
    odev = distance.device
    d = _c.to_device(distance).contiguous()
    dev = d.device
    if d.dim() != 2:
        raise ValueError("distance should be an NxM matrix")
    n, m = d.shape
    thr = as_t(thresholds, torch.float32)
    single = thr.dim() == 1
    thr = thr.reshape(1, -1) if single else thr
    scores = as_t(src_scores, torch.float64)
    st, dt_ = as_t(src_tags, torch.int32).to(dev).contiguous(), as_t(dst_tags, torch.int32).to(dev).contiguous()
    if len(scores) != n or len(st) != n or len(dt_) != m:
        raise ValueError("scores / tags do not match the distance matrix")
    order = torch.flip(torch.argsort(scores.cpu(), stable=True), dims=[0]).to(torch.int32).to(dev).contiguous()   # np.flip(np.argsort(scores))
    thr = thr.to(dev).contiguous()
    T, ncat = thr.shape
    sa = torch.empty((T, n), dtype=torch.int32, device=dev)
    da = torch.empty((T, m), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        stt = _c.match_greedy(_c.ptr(d), n, m, d.stride(0) if n else max(m, 1), _c.ptr(order), _c.ptr(st), _c.ptr(dt_), _c.ptr(thr), T, ncat, _c.ptr(sa), _c.ptr(da),
                              _c.stream_ptr())
    _c.check(stt, "match_greedy")
    if single:
        sa, da = sa[0], da[0]
    if odev.type != "cuda" or convert_numpy:
        sa, da = sa.cpu(), da.cpu()
    return (sa.numpy(), da.numpy()) if convert_numpy else (sa, da)
