#!/bin/bash
tag=${1:-x}
timeout 900 python -m pytest tests -m gpu -x -q -k "voxel" 2>&1 | tail -3
for r in 1 3 7 15; do
D3D_B200_VOX_ROLES=$r python tools/vox_probe.py 128 20 2>&1 | tail -1 | sed "s/^/roles=$r /" | tee -a gpurun_out/vox_roles_$tag.txt
done
D3D_B200_VOX_CF=8 python tools/vox_probe.py 128 20 2>&1 | tail -1 | sed "s/^/cf8 /" | tee -a gpurun_out/vox_roles_$tag.txt
D3D_B200_VOX_CF=32 python tools/vox_probe.py 128 20 2>&1 | tail -1 | sed "s/^/cf32 /" | tee -a gpurun_out/vox_roles_$tag.txt
python tools/vox_probe.py 1 20 2>&1 | tail -1 | tee -a gpurun_out/vox_roles_$tag.txt
