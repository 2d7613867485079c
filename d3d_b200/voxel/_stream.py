"""Host-side streaming of a batch of HOST frames through the GPU voxelizer.

The reference voxelizer takes one CPU tensor and returns CPU tensors (d3d/voxel/__init__.py:79-104).  For a
batch of host frames the PCIe link, not the kernel, is the bottleneck (16 B/point in, ~58 B/point out against
a kernel that runs at TB/s), so the batch is cut into chunks that are pipelined over two CUDA streams:
host->device copies of chunk c+1 and device->host copies of chunk c-1 overlap the kernel of chunk c, and the
link is used in both directions at once.  Because the C ABI packs the sparse outputs across frames, a chunk's
results come back with one contiguous copy per output array.
"""
import torch

from . import Dict, VoxelBatch

CHUNK_POINTS = 2_000_000   # ~16 KITTI frames: enough frames per launch to fill the GPU, small enough to pipeline
NSTREAMS = 2
LAZY_POINTS = True        # host batches: `points` of a frame is gathered from the caller's input on first read instead of being copied back
TIMELINE = None   # tuning: set to a list to collect (label, chunk, event) marks of one call


def _chunks(offs_host, nframes):
    out, f0 = [], 0
    while f0 < nframes:
        f1 = f0 + 1
        while f1 < nframes and int(offs_host[f1 + 1] - offs_host[f0]) <= CHUNK_POINTS:
            f1 += 1
        out.append((f0, f1))
        f0 = f1
    return out


def _mark(label, ci, stream):
    if TIMELINE is not None:
        import time
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(stream)
        TIMELINE.append((label, ci, ev, time.perf_counter()))


class _Arena:
    """One pinned host allocation per call, carved into the result arrays of every chunk.  Sized from the
    capacities (a chunk cannot keep more points or voxels than it has points), so it is requested before any
    result size is known and a steady-state caller gets the same cached block back from torch's pinned-memory
    allocator on every call."""

    def __init__(self, nbytes):
        self.buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, pin_memory=True)
        self.off = 0

    def take(self, shape, dtype):
        n = 1
        for d in shape:
            n *= int(d)
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        start = (self.off + 255) // 256 * 256
        self.off = start + nbytes
        return self.buf[start:start + nbytes].view(dtype).view(*shape)


class _Chunked:
    """The per-frame results of all chunks as one list-like object.  Frame dicts are created when they are asked for
    (a batch of 128 frames is 640 tensor views: building them eagerly costs more host time than a chunk's kernel)."""

    def __init__(self, parts):
        self.parts, self.starts, n = parts, [], 0
        for p in parts:
            self.starts.append(n)
            n += len(p)
        self.n = n

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(self.n))]
        if i < 0:
            i += self.n
        if not 0 <= i < self.n:
            raise IndexError(i)
        import bisect
        c = bisect.bisect_right(self.starts, i) - 1
        return self.parts[c][i - self.starts[c]]

    def __iter__(self):
        for p in self.parts:
            yield from p


def _slots(gen, nslots, cap_points, cap_frames, max_frame, nfeat, dev):
    """per-stream device buffers (input points, packed outputs, scratch), kept on the generator between calls:
    allocating them per chunk costs more host time than the chunk's kernel takes"""
    key = (nslots, nfeat, dev, gen._dense)
    c = getattr(gen, "_stream_slots", None)
    if c is None or c["key"] != key or c["cap_points"] < cap_points or c["cap_frames"] < cap_frames or c["max_frame"] < max_frame:
        cap_points = max(cap_points, c["cap_points"] if c and c["key"] == key else 0)
        cap_frames = max(cap_frames, c["cap_frames"] if c and c["key"] == key else 0)
        max_frame = max(max_frame, c["max_frame"] if c and c["key"] == key else 0)
        slots = []
        for _ in range(nslots):
            bufs, rows, ws = gen._alloc(cap_points, nfeat, cap_frames, max_frame, dev)
            slots.append(dict(pts=torch.empty((max(cap_points, 1), nfeat), dtype=torch.float32, device=dev), bufs=bufs, rows=rows, ws=ws,
                              offs=torch.empty(cap_frames + 1, dtype=torch.int64, device=dev)))
        c = dict(key=key, cap_points=cap_points, cap_frames=cap_frames, max_frame=max_frame, slots=slots)
        gen._stream_slots = c
    return c["slots"]


def host_batch(gen, frames, offs_host):
    """frames: list of host float32 [N_f, C] tensors (pinned memory makes the copies asynchronous).
    Returns a list-like of per-frame dicts of host tensors."""
    from .. import _cabi as _c
    _c.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    nframes = len(frames)
    nfeat = int(frames[0].shape[1])
    total = int(offs_host[-1])
    chunks = _chunks(offs_host, nframes)
    lens = offs_host[1:] - offs_host[:-1]
    max_frame = int(lens.max())
    cap_points = max(int(offs_host[f1] - offs_host[f0]) for f0, f1 in chunks)
    cap_frames = max(f1 - f0 for f0, f1 in chunks)
    nstreams = min(NSTREAMS, len(chunks))
    slots = _slots(gen, nstreams, cap_points, cap_frames, max_frame, nfeat, dev)
    streams = getattr(gen, "_stream_pool", None)
    if streams is None or len(streams) < nstreams or streams[0].device != dev:
        streams = gen._stream_pool = [torch.cuda.Stream(dev) for _ in range(nstreams)]
    streams = streams[:nstreams]
    cur = torch.cuda.current_stream(dev)
    for s in streams:
        s.wait_stream(cur)
    pending, results = [], [None] * len(chunks)
    if gen._dense:
        V, P = gen._max_voxels, gen._max_points
        per_frame = V * (P * nfeat * 4 + 24 + P + 4 + nfeat * 4) + 6 * 256
        arena = _Arena(nframes * per_frame + 24 * (nframes + len(chunks)) + 512 * len(chunks))
    else:
        arena = _Arena(total * (4 * nfeat + 8 + 8 + 4 + 24) + len(chunks) * (8 * 256 + 32) + 24 * (nframes + 1))

    def launch(ci):
        f0, f1 = chunks[ci]
        s, slot = streams[ci % nstreams], slots[ci % nstreams]
        with torch.cuda.stream(s):
            b0 = int(offs_host[f0])
            n = int(offs_host[f1]) - b0
            pts = slot["pts"][:n]
            _mark("h2d0", ci, s)
            for f in range(f0, f1):
                lo, hi = int(offs_host[f]) - b0, int(offs_host[f + 1]) - b0
                if hi > lo:
                    pts[lo:hi].copy_(frames[f], non_blocking=True)
            _mark("h2d1", ci, s)
            offs_pin = arena.take((f1 - f0 + 1,), torch.int64)     # pinned: a pageable source would block the host on the stream
            torch.sub(offs_host[f0:f1 + 1], b0, out=offs_pin)
            offs_dev = slot["offs"][:f1 - f0 + 1]
            offs_dev.copy_(offs_pin, non_blocking=True)
            res = gen._launch(pts, offs_dev, f1 - f0, int(lens[f0:f1].max()), slot["bufs"], slot["rows"], slot["ws"])
            _mark("kern1", ci, s)
            rows_h = arena.take(res.rows.shape, torch.int64)
            rows_h.copy_(res.rows, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(s)
        return res, rows_h, ev, s

    def collect(ci, item):
        res, rows_h, ev, s = item
        ev.synchronize()                       # sizes of this chunk are on the host now
        with torch.cuda.stream(s):
            if res.dense:
                # regular [frames, max_voxels, ...] arrays: copy the used prefix of every frame
                per = []
                for j in range(res.nframes):
                    nv = int(rows_h[j, 1])
                    d = Dict()
                    for k, v in res.packed.items():
                        t = arena.take((nv,) + tuple(v.shape[2:]), v.dtype)
                        t.copy_(v[j, :nv], non_blocking=True)
                        d[k] = t.view(torch.bool) if k == "voxel_pmask" else t
                    per.append(d)
                results[ci] = ("dense", per)
                return
            nk, nv = int(rows_h[-1, 0]), int(rows_h[-1, 1])
            host = {}
            _mark("d2h0", ci, s)
            for k, v in res.packed.items():
                if k == "points" and LAZY_POINTS:
                    continue   # points = input[points_mask]: 16 of the ~58 B per point that would cross the link a second time
                m = nk if k in VoxelBatch.POINT_KEYS else nv
                t = arena.take((m,) + tuple(v.shape[1:]), v.dtype)
                t.copy_(v[:m], non_blocking=True)
                host[k] = t
            _mark("d2h1", ci, s)
            f0 = chunks[ci][0]
            results[ci] = ("sparse", VoxelBatch(host, None, res.nframes, dense=False, rows_host=rows_h,
                                                lazy_points=(lambda j, f0=f0: frames[f0 + j]) if LAZY_POINTS else None))

    for ci in range(len(chunks)):
        if len(pending) == nstreams:
            collect(*pending.pop(0))
        pending.append((ci, launch(ci)))
    for it in pending:
        collect(*it)
    for s in streams:
        s.synchronize()
        cur.wait_stream(s)
    return _Chunked([r for _, r in results])
