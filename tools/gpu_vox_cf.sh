#!/bin/bash
tag=${1:-x}
timeout 900 python -m pytest tests -m gpu -x -q -k "voxel" 2>&1 | tail -3
for cf in 8 16 32 64 128; do
D3D_B200_VOX_CF=$cf python tools/vox_probe.py 128 20 2>&1 | tail -1 | sed "s/^/cf=$cf /" | tee -a gpurun_out/vox_cf_$tag.txt
done
python tools/vox_probe.py 1 20 2>&1 | tail -1 | tee -a gpurun_out/vox_cf_$tag.txt
python tools/vox_probe.py 8 20 2>&1 | tail -1 | tee -a gpurun_out/vox_cf_$tag.txt
