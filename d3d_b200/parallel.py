"""d3d_b200.parallel -- partitioning of the hot path across the GPUs of one box (one process per GPU).

The reference has no multi-device code at all; this is the new capability north_star asks
for.  The three operators shard without any data-path collective:
  * pairwise IoU: row-blocks of boxes1 (every rank holds all of boxes2, 2 MB at 100k boxes);
  * NMS and voxelization: whole frames (rank r takes frames r, r+W, r+2W, ...).
A collective appears only at the end, to hand small per-frame results (keep masks, counts, reduced IoU
products) to the consumer: `gather_ragged` is an all_gather of sizes followed by a padded all_gather,
NCCL over NVLink on the GPUs, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def row_block(n, rank, world_size):
    """[lo, hi) of the rows owned by `rank`: contiguous, balanced to within one row."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def frame_shard(nframes, rank, world_size):
    """frames owned by `rank` (round robin keeps ragged frame sizes balanced)."""
    return list(range(rank, nframes, world_size))


def iou_row_block(boxes1, boxes2, method="rbox", precise=False, rank=None, world_size=None):
    """This rank's [rows, M] slab of the N x M IoU matrix and its row range."""
    from .box import box2d_iou
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    lo, hi = row_block(boxes1.shape[0], rank, world_size)
    return box2d_iou(boxes1[lo:hi], boxes2, method=method, precise=precise), (lo, hi)


def _all_gather_flat(x):
    """all_gather of equally sized 1-D tensors into one [world * len] tensor (one collective, no per-rank Python objects)."""
    _, w = world()
    out = torch.empty(w * x.numel(), dtype=x.dtype, device=x.device)
    try:
        dist.all_gather_into_tensor(out, x.contiguous())
    except (RuntimeError, NotImplementedError, AttributeError):   # a backend without the flat form
        parts = [torch.empty_like(x) for _ in range(w)]
        dist.all_gather(parts, x.contiguous())
        out = torch.cat(parts)
    return out


def _gather_padded(t, sizes):
    """payload of every rank, given everybody's sizes (host ints): one padded all_gather, sliced per rank"""
    _, w = world()
    mx = max(max(sizes), 1)
    pad = torch.zeros(mx, dtype=t.dtype, device=t.device)
    pad[:t.numel()] = t.reshape(-1)
    buf = _all_gather_flat(pad)
    return [buf[r * mx:r * mx + sizes[r]] for r in range(w)]


def gather_ragged(t, dst=None):
    """Gather 1-D tensors of different lengths from every rank (all ranks get the list, or only `dst`).

    Sizes travel first (one all_gather of an int64 per rank, read back with a single device-to-host copy), then the payload padded to
    the longest -- the only collective the pipeline needs, run after the timed compute."""
    rank, w = world()
    if w == 1:
        return [t]
    n = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
    sizes = _all_gather_flat(n).tolist()
    out = _gather_padded(t, sizes)
    if dst is not None and rank != dst:
        return None
    return out


def gather_frames(per_frame, nframes, dtype=None, device=None):
    """Reassemble per-frame 1-D results computed under `frame_shard` into frame order on every rank.

    `dtype` / `device` of the payload must be the same on every rank; a rank whose shard is empty (more ranks than frames) cannot read
    them off its tensors, so pass them whenever that can happen (default: those of the rank's first tensor, else float32 on the
    current CUDA device under NCCL / the CPU under gloo).

    Two collectives and one device-to-host copy whatever the number of frames: the frame lengths (every rank owns at most
    ceil(nframes / world) frames, so the length vectors have one known size) and the concatenated payload."""
    rank, w = world()
    if w == 1:
        return list(per_frame)
    if device is None:
        device = per_frame[0].device if per_frame else (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu"))
    if dtype is None:
        dtype = per_frame[0].dtype if per_frame else torch.float32
    maxf = (nframes + w - 1) // w
    mine = [int(x.numel()) for x in per_frame]
    lens = torch.tensor(mine + [0] * (maxf - len(mine)), dtype=torch.int64).to(device)
    all_lens = _all_gather_flat(lens).reshape(w, maxf).tolist() if maxf else [[] for _ in range(w)]
    flat = torch.cat([x.reshape(-1) for x in per_frame]).to(device=device, dtype=dtype) if per_frame else torch.zeros(0, dtype=dtype, device=device)
    all_flat = _gather_padded(flat, [sum(row) for row in all_lens])
    out = [None] * nframes
    for r in range(w):
        off = 0
        for k, f in enumerate(frame_shard(nframes, r, w)):
            ln = all_lens[r][k]
            out[f] = all_flat[r][off:off + ln]
            off += ln
    return out
