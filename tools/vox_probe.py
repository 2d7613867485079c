#!/usr/bin/env python
"""Tuning probe (not part of the product or the tests): device-resident C2 batch through VoxelGenerator.batch_packed,
CUDA-event timed.  usage: [D3D_B200_LIB=path] python tools/vox_probe.py [frames] [steps] [algo]"""
import os
import sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import lidar, C2_BOUNDS, C2_SHAPE, C2_KW  # noqa: E402
from d3d_b200.voxel import VoxelGenerator  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 128
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
base = [lidar(100 + i) for i in range(min(F, 16))]
frames = [base[i % len(base)] for i in range(F)]
pts = torch.from_numpy(np.concatenate(frames, 0)).cuda()
offs = torch.zeros(F + 1, dtype=torch.int64)
offs[1:] = torch.tensor([len(f) for f in frames]).cumsum(0)
gen = VoxelGenerator(C2_BOUNDS, C2_SHAPE, **C2_KW)
if len(sys.argv) > 3:
    gen.algo = sys.argv[3]
for _ in range(3):
    r = gen.batch_packed(pts, offs)
torch.cuda.synchronize()
ts = []
for _ in range(steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = gen.batch_packed(pts, offs); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
rows = r.rows_host()
N = int(pts.shape[0]); K, V = int(rows[-1, 0]), int(rows[-1, 1])
ms = float(np.median(ts))
print(f"lib={os.environ.get('D3D_B200_LIB', 'default')} frames={F} N={N} K={K} V={V} median {ms:.3f} ms min {min(ts):.3f} ms "
      f"{N / ms / 1e6:.2f} Gpts/s  alg {(16 * N + 32 * K + 28 * V) / ms / 1e6:.0f} GB/s")
