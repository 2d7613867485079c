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


class _Lazy:
    """a dict value that is computed the first time it is read (d3d_b200.voxel.Dict)"""
    __slots__ = ("fn",)

    def __init__(self, fn):
        self.fn = fn


class Dict(dict):
    """attribute-access dict, standing in for addict.Dict (reference d3d/voxel/__init__.py:1).  A value may be a `_Lazy` thunk: it is
    evaluated, and replaced by its result, the first time the key is read (used for the `points` of host-streamed batches, which
    are `input[points_mask]` and need not cross the PCIe link twice)."""
    __setattr__ = dict.__setitem__

    def __getitem__(self, k):
        v = dict.__getitem__(self, k)
        if isinstance(v, _Lazy):
            v = v.fn()
            dict.__setitem__(self, k, v)
        return v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def get(self, k, default=None):
        return self[k] if k in self else default

    def values(self):
        return [self[k] for k in self.keys()]

    def items(self):
        return [(k, self[k]) for k in self.keys()]


class VoxelGenerator:
    '''
    Convert point cloud to voxels (same constructor and result keys as the reference)
    '''
    default_algo = "auto"   # back end of new instances: "auto" | "sort" | "cluster" | "tiles" | "auto_no_tiles" (results are identical)

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
        # execution back end, no effect on results: "auto" (tile pipeline, else cluster-per-frame hash path, else the sort
        # pipeline -- the first that supports the configuration), "sort", "cluster" / "tiles" (raise NotImplementedError if
        # unsupported), "auto_no_tiles" (auto without the tile pipeline)
        self.algo = VoxelGenerator.default_algo

    def __call__(self, points):
        '''
        :param points: point cloud f32[N, C]; the first three columns are xyz.  CPU or CUDA tensor.
        :return: dict with the reference's keys (points, points_mask, points_mapping, voxel_npoints, coords
                 for the sparse path; voxels, coords, voxel_pmask, voxel_npoints[, aggregates] when dense)
        '''
        return self.batch([points])[0]

    def batch(self, frames):
        """Voxelize a list of frames in one launch; returns a sequence with one result dict per frame.
        CUDA frames: results stay on the device.  Host frames: the batch is streamed through the GPU
        (pinned staging, copies overlapped with compute) and the results come back as host tensors."""
        if len(frames) == 0:
            return []
        odev = frames[0].device
        nfeat = frames[0].shape[1]
        for f in frames:
            if f.dtype != torch.float32:
                raise RuntimeError("expected scalar type Float")  # reference: accessor<float,2> type error
            if f.dim() != 2 or f.shape[1] != nfeat:
                raise ValueError("all frames must be [N, C] with the same C")
        lens = torch.tensor([int(f.shape[0]) for f in frames], dtype=torch.int64)
        offs_host = torch.zeros(len(frames) + 1, dtype=torch.int64)
        offs_host[1:] = lens.cumsum(0)
        if odev.type != "cuda":
            from ._stream import host_batch
            return host_batch(self, frames, offs_host)
        pts = frames[0].contiguous() if len(frames) == 1 else torch.cat(list(frames), 0)
        return self._run(pts, offs_host)

    def batch_packed(self, points, offsets_host, offsets_dev=None):
        """Batch form without the concatenation: `points` f32[total, C] already holds the frames back to
        back (CUDA tensor), `offsets_host` is the int64[nframes+1] CPU tensor of frame boundaries
        (`offsets_dev`: the same offsets already on the device, saves the copy)."""
        if points.dtype != torch.float32:
            raise RuntimeError("expected scalar type Float")
        return self._run(_c.to_device(points), offsets_host.to(torch.int64), offs_dev=offsets_dev)

    def _alloc(self, total, nfeat, nframes, max_frame_points, dev):
        """device buffers of one launch: packed outputs, per-frame row table, scratch"""
        tcap = max(total, 1)
        if self._dense:
            V, P = self._max_voxels, self._max_points
            rows = torch.empty((nframes, 2), dtype=torch.int64, device=dev)
            bufs = dict(voxels=torch.empty((nframes, V, P, nfeat), dtype=torch.float32, device=dev),
                        coords=torch.empty((nframes, V, 3), dtype=torch.int64, device=dev),
                        voxel_pmask=torch.empty((nframes, V, P), dtype=torch.uint8, device=dev),
                        voxel_npoints=torch.empty((nframes, V), dtype=torch.int32, device=dev))
            if self._reduction != ReductionType.NONE:
                bufs["aggregates"] = torch.empty((nframes, V, nfeat), dtype=torch.float32, device=dev)
        else:
            rows = torch.empty((nframes + 1, 2), dtype=torch.int64, device=dev)
            bufs = dict(points=torch.empty((tcap, nfeat), dtype=torch.float32, device=dev),
                        points_mask=torch.empty(tcap, dtype=torch.int64, device=dev),
                        points_mapping=torch.empty(tcap, dtype=torch.int64, device=dev),
                        voxel_npoints=torch.empty(tcap, dtype=torch.int32, device=dev),
                        coords=torch.empty((tcap, 3), dtype=torch.int64, device=dev))
        ws = _c.workspace(_c.voxelize_workspace_bytes(total, nframes, max_frame_points), dev)
        return bufs, rows, ws

    def _launch(self, pts, offs, nframes, max_frame_points, bufs, rows, ws):
        """one C-ABI call on the current stream; bufs/rows/ws may be larger than this launch needs"""
        total, nfeat = int(pts.shape[0]), int(pts.shape[1])
        p = _c.VoxelParams.from_buffer_copy(self._params)
        p.max_frame_points = max_frame_points
        p.algo = {"auto": 0, "sort": 1, "cluster": 2, "tiles": 3, "auto_no_tiles": 4}[self.algo]
        with torch.cuda.device(pts.device):
            if self._dense:
                if nframes < bufs["voxels"].shape[0]:
                    bufs = {k: v[:nframes] for k, v in bufs.items()}
                    rows = rows[:nframes]
                st = _c.voxelize_dense(_c.ptr(pts), total, nfeat, _c.ptr(offs), nframes, C.byref(p), _c.ptr(bufs["voxels"]), _c.ptr(bufs["coords"]),
                                       _c.ptr(bufs["voxel_pmask"]), _c.ptr(bufs["voxel_npoints"]), _c.ptr(bufs.get("aggregates")), _c.ptr(rows),
                                       _c.ptr(ws), ws.numel(), _c.stream_ptr())
                _c.check(st, "voxelize_3d_dense")
                return VoxelBatch(bufs, rows, nframes, dense=True)
            st = _c.voxelize_sparse(_c.ptr(pts), total, nfeat, _c.ptr(offs), nframes, C.byref(p), _c.ptr(bufs["points"]), _c.ptr(bufs["points_mask"]),
                                    _c.ptr(bufs["points_mapping"]), _c.ptr(bufs["voxel_npoints"]), _c.ptr(bufs["coords"]), _c.ptr(rows),
                                    _c.ptr(ws), ws.numel(), _c.stream_ptr())
            _c.check(st, "voxelize_3d_sparse/filter")
            return VoxelBatch(bufs, rows[:nframes + 1], nframes, dense=False)

    def _run(self, pts, offs_host, offs_dev=None):
        """launch on the current stream; nothing is read back until a frame of the result is indexed"""
        dev = pts.device
        total, nfeat = int(pts.shape[0]), int(pts.shape[1])
        nframes = offs_host.numel() - 1
        offs = offs_host.to(dev, non_blocking=True) if offs_dev is None else offs_dev
        max_frame = int((offs_host[1:] - offs_host[:-1]).max()) if nframes > 0 else 0
        with torch.cuda.device(dev):
            bufs, rows, ws = self._alloc(total, nfeat, nframes, max_frame, dev)
        return self._launch(pts, offs, nframes, max_frame, bufs, rows, ws)


class VoxelBatch:
    """Result of a batched voxelization: behaves like a list of per-frame result dicts.

    The per-frame tensors are views into packed buffers (`.packed`); the per-frame sizes live on the device
    (`.rows`) and are read back -- the only synchronisation -- the first time a frame is indexed."""
    POINT_KEYS = ("points", "points_mask", "points_mapping")

    def __init__(self, bufs, rows, nframes, dense, rows_host=None, lazy_points=None):
        self.packed, self.rows, self.nframes, self.dense = bufs, rows, nframes, dense
        self._rows_host = rows_host
        self._lazy_points = lazy_points   # frame -> input tensor: `points` of a frame is input[points_mask], built on first read
        self._cache = {}

    def __len__(self):
        return self.nframes

    def rows_host(self):
        if self._rows_host is None:
            self._rows_host = self.rows.cpu()
        return self._rows_host

    def __getitem__(self, f):
        if isinstance(f, slice):
            return [self[i] for i in range(*f.indices(self.nframes))]
        if f < 0:
            f += self.nframes
        if not 0 <= f < self.nframes:
            raise IndexError(f)
        r = self._cache.get(f)
        if r is None:
            rh = self.rows_host()
            if self.dense:
                nv = int(rh[f, 1])
                r = Dict((k, v[f, :nv]) for k, v in self.packed.items())
                r["voxel_pmask"] = r["voxel_pmask"].view(torch.bool)
            else:
                k0, k1, v0, v1 = int(rh[f, 0]), int(rh[f + 1, 0]), int(rh[f, 1]), int(rh[f + 1, 1])
                r = Dict((k, v[k0:k1] if k in self.POINT_KEYS else v[v0:v1]) for k, v in self.packed.items())
                if self._lazy_points is not None and "points" not in r:
                    src, mask = self._lazy_points(f), r["points_mask"]
                    dict.__setitem__(r, "points", _Lazy(lambda src=src, mask=mask: src.index_select(0, mask)))
            self._cache[f] = r
        return r

    def __iter__(self):
        return (self[i] for i in range(self.nframes))
