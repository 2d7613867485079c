"""d3d_b200.voxel -- point-cloud voxelization.  Mirrors reference d3d/voxel/__init__.py:12-104
(VoxelGenerator) and the enums of d3d/voxel/voxelize.h:5-7 on top of the batched C ABI.

New capability over the reference (which voxelizes one CPU tensor at a time): `VoxelGenerator.batch`
takes a list of frames and voxelizes them in one launch sequence; `__call__` is the one-frame case."""
import ctypes as C
import enum

import torch

from .. import _cabi as _c


class ReductionType(enum.IntEnum):
    NONE = 0
    MEAN = 1
    MAX = 2
    MIN = 3


class MaxPointsFilterType(enum.IntEnum):
    NONE = 0
    TRIM = 1
    FARTHEST_SAMPLING = 2


class MaxVoxelsFilterType(enum.IntEnum):
    NONE = 0
    TRIM = 1
    DESCENDING = 2


class Dict(dict):
    """attribute-access dict, standing in for addict.Dict (reference d3d/voxel/__init__.py:1)"""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


class VoxelGenerator:
    '''
    Convert point cloud to voxels (same constructor and result keys as the reference)
    '''
    default_algo = "auto"   # back end of new instances: "auto" | "sort" | "cluster" (results are identical)

    def __init__(self, bounds, shape,
        min_points=0, max_points=30, max_voxels=20000,
        max_points_filter=None, max_voxels_filter=None,
        reduction=None, dense=False):
        self._bounds = torch.tensor(bounds, dtype=torch.float)
        self._shape = torch.tensor(shape, dtype=torch.int32)
        self._min_points = min_points
        self._max_points = max_points
        self._max_voxels = max_voxels
        self._dense = dense

        # same fp32 tensor arithmetic as the reference (d3d/voxel/__init__.py:39-46): the voxel size is
        # (hi-lo)/shape rounded to float, NOT the literal size
        bounds_array = self._bounds.reshape(3, 2)
        self._size = (bounds_array[:, 1] - bounds_array[:, 0]) / self._shape
        bounds_dist = bounds_array[:, 0] / self._size
        if torch.any(torch.abs(torch.round(bounds_dist) - bounds_dist) > 1e-3):
            raise ValueError("The voxelization grids is not aligned with the origin, which could lead to unexpected behavior!")
        self._offset = torch.round(bounds_dist).int()
        self._vbounds = torch.round(bounds_array / self._size.reshape(3, 1)).long()

        reduction = (reduction or "NONE").upper()
        if reduction != "NONE" and not dense:
            raise ValueError("Reduction is only for dense voxelization!")
        if reduction in ReductionType.__members__:
            self._reduction = ReductionType[reduction]
        else:
            raise ValueError("Unsupported reduction type in VoxelGenerator!")

        max_points_filter = (max_points_filter or "NONE").upper()
        if max_points_filter in MaxPointsFilterType.__members__:
            self._max_points_filter = MaxPointsFilterType[max_points_filter]
        else:
            raise ValueError("Unsupported maximum points filter in VoxelGenerator!")

        max_voxels_filter = (max_voxels_filter or "NONE").upper()
        if max_voxels_filter in MaxVoxelsFilterType.__members__:
            self._max_voxels_filter = MaxVoxelsFilterType[max_voxels_filter]
        else:
            raise ValueError("Unsupported maximum voxels filter in VoxelGenerator!")

        if dense:
            if min_points > 0:
                raise NotImplementedError("Minimum points filtering is not implemented for dense")
            if self._max_points_filter not in [MaxPointsFilterType.NONE, MaxPointsFilterType.TRIM]:
                raise NotImplementedError("Only trim is implemented for max points filtering")
            if self._max_voxels_filter not in [MaxVoxelsFilterType.NONE, MaxVoxelsFilterType.TRIM]:
                raise NotImplementedError("Only trim is implemented for max voxels filtering")

        p = _c.VoxelParams()
        for d in range(3):
            p.size[d] = float(self._size[d])
            p.vlo[d] = int(self._vbounds[d, 0])
            p.vhi[d] = int(self._vbounds[d, 1])
            p.offset[d] = int(self._offset[d])
            p.shape[d] = int(self._shape[d])
        for k in range(6):
            p.bound[k] = float(self._bounds[k])
        p.min_points, p.max_points, p.max_voxels = int(min_points), int(max_points), int(max_voxels)
        p.max_points_filter = int(self._max_points_filter)
        p.max_voxels_filter = int(self._max_voxels_filter)
        p.reduction = int(self._reduction)
        self._params = p
        # execution back end, no effect on results: "auto" (cluster-per-frame hash path when it supports the
        # configuration, else the sort pipeline), "sort", "cluster" (raises NotImplementedError if unsupported)
        self.algo = VoxelGenerator.default_algo

    def __call__(self, points):
        '''
        :param points: point cloud f32[N, C]; the first three columns are xyz.  CPU or CUDA tensor.
        :return: dict with the reference's keys (points, points_mask, points_mapping, voxel_npoints, coords
                 for the sparse path; voxels, coords, voxel_pmask, voxel_npoints[, aggregates] when dense)
        '''
        return self.batch([points])[0]

    def batch(self, frames):
        """Voxelize a list of frames in one launch sequence; returns one dict per frame."""
        if len(frames) == 0:
            return []
        odev = frames[0].device
        nfeat = frames[0].shape[1]
        for f in frames:
            if f.dtype != torch.float32:
                raise RuntimeError("expected scalar type Float")  # reference: accessor<float,2> type error
            if f.dim() != 2 or f.shape[1] != nfeat:
                raise ValueError("all frames must be [N, C] with the same C")
        lens = [int(f.shape[0]) for f in frames]
        pts = _c.to_device(frames[0]) if len(frames) == 1 else torch.cat([_c.to_device(f) for f in frames], 0)
        offs_host = torch.zeros(len(frames) + 1, dtype=torch.int64)
        offs_host[1:] = torch.tensor(lens, dtype=torch.int64).cumsum(0)
        return self._run(pts, offs_host, odev)

    def batch_packed(self, points, offsets_host):
        """Batch form without the concatenation: `points` f32[total, C] already holds the frames back to
        back (CUDA tensor), `offsets_host` is the int64[nframes+1] CPU tensor of frame boundaries."""
        if points.dtype != torch.float32:
            raise RuntimeError("expected scalar type Float")
        return self._run(_c.to_device(points), offsets_host.to(torch.int64), points.device)

    def _run(self, pts, offs_host, odev):
        dev = pts.device
        total, nfeat = int(pts.shape[0]), int(pts.shape[1])
        nframes = offs_host.numel() - 1
        offs = offs_host.to(dev, non_blocking=True)
        counts = torch.empty((nframes, 2), dtype=torch.int64, device=dev)
        p = _c.VoxelParams.from_buffer_copy(self._params)
        p.max_frame_points = int((offs_host[1:] - offs_host[:-1]).max()) if nframes > 0 else 0
        p.algo = {"auto": 0, "sort": 1, "cluster": 2}[self.algo]
        ws = _c.workspace(_c.voxelize_workspace_bytes(total, nframes, p.max_frame_points), dev)
        tcap = max(total, 1)
        with torch.cuda.device(dev):
            if self._dense:
                V, P = self._max_voxels, self._max_points
                voxels = torch.empty((nframes, V, P, nfeat), dtype=torch.float32, device=dev)
                coords = torch.empty((nframes, V, 3), dtype=torch.int64, device=dev)
                pmask = torch.empty((nframes, V, P), dtype=torch.uint8, device=dev)
                npts = torch.empty((nframes, V), dtype=torch.int32, device=dev)
                aggr = torch.empty((nframes, V, nfeat), dtype=torch.float32, device=dev) if self._reduction != ReductionType.NONE else None
                st = _c.voxelize_dense(_c.ptr(pts), total, nfeat, _c.ptr(offs), nframes, C.byref(p), _c.ptr(voxels), _c.ptr(coords),
                                       _c.ptr(pmask), _c.ptr(npts), _c.ptr(aggr), _c.ptr(counts), _c.ptr(ws), ws.numel(), _c.stream_ptr())
                _c.check(st, "voxelize_3d_dense")
                cnt = counts.cpu()   # the one readback: per-frame sizes
                out = []
                for f in range(nframes):
                    nv = int(cnt[f, 1])
                    r = Dict(voxels=voxels[f, :nv], coords=coords[f, :nv], voxel_pmask=pmask[f, :nv].view(torch.bool),
                             voxel_npoints=npts[f, :nv])
                    if aggr is not None:
                        r["aggregates"] = aggr[f, :nv]
                    out.append(self._home(r, odev))
                return out
            out_points = torch.empty((tcap, nfeat), dtype=torch.float32, device=dev)
            out_mask = torch.empty(tcap, dtype=torch.int64, device=dev)
            out_map = torch.empty(tcap, dtype=torch.int64, device=dev)
            out_np = torch.empty(tcap, dtype=torch.int32, device=dev)
            out_co = torch.empty((tcap, 3), dtype=torch.int64, device=dev)
            st = _c.voxelize_sparse(_c.ptr(pts), total, nfeat, _c.ptr(offs), nframes, C.byref(p), _c.ptr(out_points), _c.ptr(out_mask),
                                    _c.ptr(out_map), _c.ptr(out_np), _c.ptr(out_co), _c.ptr(counts), _c.ptr(ws), ws.numel(), _c.stream_ptr())
            _c.check(st, "voxelize_3d_sparse/filter")
            cnt = counts.cpu()
            out = []
            for f in range(nframes):
                b = int(offs_host[f]); k = int(cnt[f, 0]); nv = int(cnt[f, 1])
                out.append(self._home(Dict(points=out_points[b:b + k], points_mask=out_mask[b:b + k], points_mapping=out_map[b:b + k],
                                           voxel_npoints=out_np[b:b + nv], coords=out_co[b:b + nv]), odev))
            return out

    @staticmethod
    def _home(r, odev):
        """results go back to where the input came from; host copies land in pinned memory and are
        queued asynchronously (one stream sync per dict instead of one per tensor)"""
        if odev.type != "cuda":
            for k in list(r.keys()):
                src = r[k]
                dst = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
                dst.copy_(src, non_blocking=True)
                r[k] = dst
            torch.cuda.current_stream().synchronize()
        return r
