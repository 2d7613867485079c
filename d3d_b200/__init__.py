"""d3d_b200 -- B200-native (sm_100a) implementation of d3d's data-parallel geometry operators.

Drop-in for the hot path of cmpute/d3d: `d3d_b200.box` (box2d_iou, box2d_nms), `d3d_b200.voxel`
(VoxelGenerator) and `d3d_b200.point` (aligned_scatter) keep the reference's call signatures, tensor
layouts and return dtypes (reference d3d/box/__init__.py, d3d/voxel/__init__.py, d3d/point/__init__.py).
All compute runs in hand-written CUDA behind the C ABI of include/d3d_b200.h; there is no CPU path:
importing the package without the built extension raises ImportError.
"""
from . import _cabi  # noqa: F401  (fails loudly when libd3d_b200.so is missing)

__all__ = ["box", "voxel", "point", "parallel"]
