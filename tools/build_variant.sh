#!/bin/bash
# tuning helper: build a variant of the library with extra nvcc flags, e.g.
#   tools/build_variant.sh timing -DD3D_VC_TIMING      -> d3d_b200/libd3d_b200_timing.so  (use with D3D_B200_LIB=...)
set -e
name=$1; shift
cd "$(dirname "$0")/../d3d_b200/csrc"
d=$(mktemp -d)
for f in core prims iou nms voxel voxel_cluster voxel_tiles scatter crop dist match iou_grad; do
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c $f.cu -o $d/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libd3d_b200_$name.so $d/*.o -lcudart
rm -rf $d
echo built d3d_b200/libd3d_b200_$name.so
