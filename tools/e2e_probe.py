#!/usr/bin/env python
"""Tuning probe: host frames -> VoxelGenerator.batch -> host results, wall-clock per step."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import lidar, C2_BOUNDS, C2_SHAPE, C2_KW  # noqa: E402
from d3d_b200.voxel import VoxelGenerator, _stream  # noqa: E402
F = int(sys.argv[1]) if len(sys.argv) > 1 else 128
_stream.NSTREAMS = int(sys.argv[2]) if len(sys.argv) > 2 else 2
_stream.CHUNK_POINTS = int(sys.argv[3]) if len(sys.argv) > 3 else 2_000_000
base = [lidar(100 + i) for i in range(min(F, 16))]
host = [torch.from_numpy(base[i % len(base)]).pin_memory() for i in range(F)]
gen = VoxelGenerator(C2_BOUNDS, C2_SHAPE, **C2_KW)
for step in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = gen.batch(host)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"streams={_stream.NSTREAMS} chunk={_stream.CHUNK_POINTS} step {step}: {(t1 - t0) * 1e3:.1f} ms  ({F * 120000 / (t1 - t0) / 1e6:.0f} Mpts/s)", flush=True)
    del r

_stream.TIMELINE = []
torch.cuda.synchronize(); t0 = time.perf_counter()
e0 = torch.cuda.Event(enable_timing=True); e0.record()
r = gen.batch(host)
torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"timeline step {(t1 - t0) * 1e3:.1f} ms")
for label, ci, ev, th in _stream.TIMELINE:
    print(f"  chunk {ci} {label:6s} gpu {e0.elapsed_time(ev):7.2f} ms   host-issued {(th - t0) * 1e3:7.2f} ms")
import cProfile, pstats
del r
_stream.TIMELINE = None
for _ in range(3):
    r = gen.batch(host); torch.cuda.synchronize(); del r
pr = cProfile.Profile(); pr.enable()
for _ in range(3):
    r = gen.batch(host); torch.cuda.synchronize(); del r
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
