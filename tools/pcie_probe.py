#!/usr/bin/env python
"""Tuning probe: pinned host <-> device copy bandwidth, one direction and both at once."""
import time, torch
n = 256 << 20
h1 = torch.empty(n, dtype=torch.uint8, pin_memory=True); h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
a = t(lambda: d1.copy_(h1, non_blocking=True)); print(f"H2D {n / a / 1e9:.1f} GB/s")
b = t(lambda: h2.copy_(d2, non_blocking=True)); print(f"D2H {n / b / 1e9:.1f} GB/s")
def both():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
c = t(both); print(f"both directions at once: {2 * n / c / 1e9:.1f} GB/s aggregate ({c * 1e3:.1f} ms for 2 x 256 MiB)")
