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


def gather_ragged(t, dst=None):
    """Gather 1-D tensors of different lengths from every rank (all ranks get the list, or only `dst`).

    Sizes travel first (all_gather of one int64), then the payload padded to the longest -- the only
    collective the pipeline needs, run after the timed compute."""
    rank, w = world()
    if w == 1:
        return [t]
    n = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(w)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    mx = max(max(sizes), 1)
    pad = torch.zeros(mx, dtype=t.dtype, device=t.device)
    pad[:t.numel()] = t.reshape(-1)
    bufs = [torch.empty_like(pad) for _ in range(w)]
    dist.all_gather(bufs, pad)
    out = [b[:s] for b, s in zip(bufs, sizes)]
    if dst is not None and rank != dst:
        return None
    return out


def gather_frames(per_frame, nframes, dtype=None, device=None):
    """Reassemble per-frame 1-D results computed under `frame_shard` into frame order on every rank.

    `dtype` / `device` of the payload must be the same on every rank; a rank whose shard is empty (more ranks than frames) cannot read
    them off its tensors, so pass them whenever that can happen (default: those of the rank's first tensor, else float32 on the
    current CUDA device under NCCL / the CPU under gloo)."""
    rank, w = world()
    if w == 1:
        return list(per_frame)
    if device is None:
        device = per_frame[0].device if per_frame else (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu"))
    if dtype is None:
        dtype = per_frame[0].dtype if per_frame else torch.float32
    lens = torch.tensor([x.numel() for x in per_frame], dtype=torch.int64, device=device)
    flat = torch.cat([x.reshape(-1) for x in per_frame]).to(device=device, dtype=dtype) if per_frame else torch.zeros(0, dtype=dtype, device=device)
    all_lens = gather_ragged(lens)
    all_flat = gather_ragged(flat)
    out = [None] * nframes
    for r in range(w):
        off = 0
        for k, f in enumerate(frame_shard(nframes, r, w)):
            ln = int(all_lens[r][k])
            out[f] = all_flat[r][off:off + ln]
            off += ln
    return out
