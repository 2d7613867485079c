#!/bin/bash
# tuning helper (run under gpurun): the C2 probe for every library variant named (no tests)
tag=${1:-x}; shift
for v in "" "$@"; do
  lib=$PWD/d3d_b200/libd3d_b200${v:+_$v}.so
  for rep in 1 2; do
    D3D_B200_LIB=$lib python tools/vox_probe.py 128 30 2>&1 | tail -1 | sed "s/^/v=${v:-default} /" | cut -c1-170 | tee -a gpurun_out/vox_ab_$tag.txt
  done
done
