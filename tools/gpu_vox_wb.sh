#!/bin/bash
# tuning helper (run under gpurun): voxel parity tests on a library variant, then the C2 probe for the default and the variants
# usage: tools/gpu_vox_wb.sh tag testvariant [variant ...]
tag=${1:-x}; tv=$2; shift; shift
D3D_B200_LIB=$PWD/d3d_b200/libd3d_b200_$tv.so timeout 900 python -m pytest tests -m gpu -x -q -k "voxel" 2>&1 | tail -3 | tee gpurun_out/vox_${tag}_tests.txt
for v in "" $tv "$@"; do
  lib=$PWD/d3d_b200/libd3d_b200${v:+_$v}.so
  for rep in 1 2; do
    D3D_B200_LIB=$lib python tools/vox_probe.py 128 30 2>&1 | tail -1 | sed "s/^/v=${v:-default} /" | cut -c1-170 | tee -a gpurun_out/vox_$tag.txt
  done
done
