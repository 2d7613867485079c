"""world_size-2 `gloo` tests of d3d_b200/parallel.py: the partitioning of the hot path over the GPUs of one
box (row-blocks for pairwise IoU, whole frames for NMS / voxelization) and the end-of-pipeline ragged gather.
Runs on CPU: the ranks compute their shard with the ORACLE (the checker), so what is tested here is that the
shards tile the problem exactly and that the gather reassembles them in order."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.timeout(180)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from d3d_b200 import parallel as P
        from oracle import oracle as O
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from conftest import gen_boxes, lidar, proposals
        # ---- pairwise IoU: row-blocks tile [0, N) and the gathered slabs equal the full matrix
        rng = np.random.default_rng(11)
        A, B = gen_boxes(rng, 101), gen_boxes(rng, 37)          # 101 rows: uneven split
        lo, hi = P.row_block(len(A), rank, world)
        slab = torch.from_numpy(O.iou2dr(A[lo:hi], B))
        parts = P.gather_ragged(slab.reshape(-1))
        full = torch.cat(parts).reshape(len(A), len(B)).numpy()
        assert np.array_equal(full, O.iou2dr(A, B))
        bounds = [P.row_block(len(A), r, world) for r in range(world)]
        assert bounds[0][0] == 0 and bounds[-1][1] == len(A) and all(bounds[i][1] == bounds[i + 1][0] for i in range(world - 1))
        # ---- NMS by frame: 5 frames of different sizes over 2 ranks, keep-index lists gathered in frame order
        frames = [proposals(np.random.default_rng(50 + f), 200 + 37 * f, 12) for f in range(5)]
        mine = P.frame_shard(len(frames), rank, world)
        keeps = [torch.from_numpy(np.nonzero(O.box2d_nms(frames[f][0], frames[f][1], "rbox", iou_threshold=0.5))[0]) for f in mine]
        allk = P.gather_frames(keeps, len(frames))
        for f, (bx, sc) in enumerate(frames):
            assert np.array_equal(allk[f].numpy(), np.nonzero(O.box2d_nms(bx, sc, "rbox", iou_threshold=0.5))[0]), f
        # ---- voxelization by frame: per-frame voxel counts and point->voxel maps
        clouds = [lidar(np.random.default_rng(70 + f), 3000 + 500 * f) for f in range(4)]
        gen = O.VoxelGenerator([0, 70.4, -40, 40, -3, 1], [1408, 1600, 40], max_points=5, max_points_filter="trim")
        mine = P.frame_shard(len(clouds), rank, world)
        maps = [torch.from_numpy(gen(clouds[f])["points_mapping"]) for f in mine]
        allm = P.gather_frames(maps, len(clouds))
        for f, c in enumerate(clouds):
            assert np.array_equal(allm[f].numpy(), gen(c)["points_mapping"]), f
        # ---- dst-only gather and the empty-shard case (more ranks than frames)
        one = P.gather_frames([torch.arange(3)] if rank == 0 else [], 1, dtype=torch.int64)
        assert len(one) == 1 and one[0].tolist() == [0, 1, 2] and one[0].dtype == torch.int64
        kb = P.gather_frames([torch.tensor([True, False, True])] if rank == 0 else [], 1, dtype=torch.bool)   # payload dtype agreed across ranks
        assert kb[0].dtype == torch.bool and kb[0].tolist() == [True, False, True]
        r = P.gather_ragged(torch.arange(rank + 1), dst=0)
        assert (r is None) == (rank != 0)
        if rank == 0:
            assert [x.tolist() for x in r] == [[0], [0, 1]]
        q.put((rank, "ok"))
    except BaseException as e:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
        raise
    finally:
        dist.destroy_process_group()


def test_partition_and_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(150)
    res = dict(q.get() for _ in range(2))
    assert res == {0: "ok", 1: "ok"}, res


def test_row_block_and_frame_shard_cover_everything():
    from d3d_b200 import parallel as P
    for n in (0, 1, 7, 100_000):
        for w in (1, 2, 4, 8):
            blocks = [P.row_block(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            assert max(b[1] - b[0] for b in blocks) - min(b[1] - b[0] for b in blocks) <= 1
            shards = [P.frame_shard(n if n < 1000 else 64, r, w) for r in range(w)]
            assert sorted(sum(shards, [])) == list(range(n if n < 1000 else 64))
