#!/usr/bin/env python
"""Debug probe: frame-batched NMS against per-frame NMS and the oracle on the frames of test_nms_batch_equals_per_frame."""
import os
import sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import proposals  # noqa: E402
from d3d_b200 import _cabi  # noqa: E402
from d3d_b200.box import box2d_nms, box2d_nms_batch  # noqa: E402
from oracle import oracle  # noqa: E402

dev = torch.device("cuda:0")
rng = np.random.default_rng(17)
sizes = (700, 0, 1, 64, 4096, 65, 1500, 333)
frames = [proposals(rng, n, max(n // 20, 1), extent=30.0) if n else (np.zeros((0, 5)), np.zeros(0)) for n in sizes]
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
b, s = frames[4]
exp = oracle.box2d_nms(b, s, "rbox", iou_threshold=0.5, score_threshold=0.0, cuda_score_rule=True)
print("oracle kept", int(exp.sum()))
for rep in range(3):
    for bpath in (None, 1):
        _cabi.tuning_set("D3D_B200_NMS_BATCH_PATH", bpath)
        keeps = box2d_nms_batch([t(x) for x, _ in frames], [t(y) for _, y in frames], iou_method="rbox", iou_threshold=0.5, score_threshold=0.0)
        k = keeps[4].cpu().numpy()
        alone = box2d_nms_batch([t(b)], [t(s)], iou_method="rbox", iou_threshold=0.5)[0].cpu().numpy()
        print("rep", rep, "batch path", bpath, "kept", int(k.sum()), "differs from oracle at", np.nonzero(k != exp)[0][:10], "| frame alone differs at", np.nonzero(alone != exp)[0][:10])
    _cabi.tuning_set("D3D_B200_NMS_BATCH_PATH", None)
    for fix in (0, 1):
        _cabi.tuning_set("D3D_B200_NMS_FIX", fix)
        k1 = box2d_nms(t(b), t(s), "rbox", iou_threshold=0.5).cpu().numpy()
        print("rep", rep, "single, fix", fix, "kept", int(k1.sum()), "differs from oracle at", np.nonzero(k1 != exp)[0][:10])
    _cabi.tuning_set("D3D_B200_NMS_FIX", None)
bad = np.nonzero(k != exp)[0]
if len(bad):
    i = bad[0]
    order = np.lexsort((np.arange(len(s)), -s))
    rank = np.empty(len(s), int); rank[order] = np.arange(len(s))
    ious = oracle.iou2dr(b[i:i + 1], b)[0]
    near = np.nonzero(ious > 0.3)[0]
    print("box", i, "rank", rank[i], "score", s[i], "oracle keep", exp[i], "batch keep", k[i])
    for j in near[np.argsort(rank[near])]:
        print("   overlaps box", j, "rank", rank[j], "iou %.12f" % ious[j], "oracle keep", exp[j], "batch keep", k[j])
