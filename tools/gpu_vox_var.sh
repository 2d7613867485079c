#!/bin/bash
tag=${1:-x}
for v in "" _vB _vC _vD _vE; do
for cf in 16 128; do
D3D_B200_LIB=$PWD/d3d_b200/libd3d_b200$v.so D3D_B200_VOX_CF=$cf python tools/vox_probe.py 128 20 2>&1 | tail -1 | sed "s/^/v=$v cf=$cf /" | cut -c1-150 | tee -a gpurun_out/vox_var_$tag.txt
done
done
