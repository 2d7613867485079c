#!/usr/bin/env python
"""Debug probe: cluster path vs sort path on the golden small cloud; prints the points whose keep decision differs."""
import os
import sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from d3d_b200.voxel import VoxelGenerator  # noqa: E402

g = np.load(os.path.join(ROOT, "tests/golden/voxel_c2small.npz"))
pts = g["points"]
kw = dict(max_points=int(sys.argv[1]) if len(sys.argv) > 1 else 5, max_points_filter="trim")
res = {}
for algo in ("sort", "cluster"):
    gen = VoxelGenerator(g["bounds"].tolist(), g["shape"].tolist(), **kw)
    gen.algo = algo
    r = gen(torch.from_numpy(pts).cuda())
    res[algo] = {k: v.cpu().numpy() for k, v in r.items()}
a, b = res["sort"], res["cluster"]
print("kept", len(a["points_mask"]), len(b["points_mask"]), "voxels", len(a["coords"]), len(b["coords"]))
miss = np.setdiff1d(a["points_mask"], b["points_mask"])
extra = np.setdiff1d(b["points_mask"], a["points_mask"])
print("missing", miss, "extra", extra)
bd = np.array(g["bounds"], dtype=np.float64).reshape(3, 2)
size = ((bd[:, 1] - bd[:, 0]) / np.array(g["shape"])).astype(np.float32)
cell = np.floor(pts[:, :3] / size).astype(np.int64)
for m in miss:
    same = np.nonzero((cell == cell[m]).all(1))[0]
    print("point", m, "cell", cell[m], "voxel members", same, "kept by cluster path:", np.intersect1d(same, b["points_mask"]))
