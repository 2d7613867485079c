#!/usr/bin/env python
"""Tuning probe (not product): the C5 step (voxelize -> batched NMS) on F frames of one GPU; device time, host time of the calls,
and each operator alone.  usage: python tools/c5_small_probe.py [frames]"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import lidar, proposals, C2_BOUNDS, C2_SHAPE, C2_KW  # noqa: E402
from d3d_b200.voxel import VoxelGenerator  # noqa: E402
from d3d_b200.box import box2d_nms_batch  # noqa: E402
F = int(sys.argv[1]) if len(sys.argv) > 1 else 8
NP = 4096
clouds = [lidar(500 + f) for f in range(F)]
props = [proposals(900 + f, NP, 160) for f in range(F)]
pts = torch.cat([torch.from_numpy(x) for x in clouds], 0).cuda()
offs = torch.zeros(F + 1, dtype=torch.int64); offs[1:] = torch.tensor([len(x) for x in clouds]).cumsum(0)
offs_dev = offs.cuda()
pb = torch.cat([torch.from_numpy(b) for b, _ in props], 0).cuda()
ps = torch.cat([torch.from_numpy(s) for _, s in props], 0).cuda()
poffs = torch.arange(F + 1, dtype=torch.int64) * NP; poffs_dev = poffs.cuda()
gen = VoxelGenerator(C2_BOUNDS, C2_SHAPE, **C2_KW)
def vox(): return gen.batch_packed(pts, offs, offs_dev)
def nms(): return box2d_nms_batch(pb, ps, poffs, iou_method="rbox", iou_threshold=0.5, offsets_dev=poffs_dev)
def both(): vox(); nms()
def run(fn, name, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    dev, host = [], []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record(); fn(); e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
        dev.append(e0.elapsed_time(e1)); host.append((t1 - t0) * 1e3)
    # back to back without synchronise (the bench loop)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"F={F} {name}: device {np.median(dev):.3f} ms, host issue {np.median(host):.3f} ms, back-to-back {e0.elapsed_time(e1) / n:.3f} ms/step")
run(vox, "voxel"); run(nms, "nms"); run(both, "both")
