#!/bin/bash
# tuning helper (run under gpurun): voxel parity tests, C2 probe timing, one ncu --set full capture of the cluster kernel
tag=${1:-x}
timeout 600 python -m pytest tests -m gpu -x -q -k "voxel" 2>&1 | tail -3
python tools/vox_probe.py 128 20 2>&1 | tail -1 | tee gpurun_out/vox_$tag.txt
python tools/vox_probe.py 144 20 2>&1 | tail -1 | tee -a gpurun_out/vox_$tag.txt
if [ "$2" != "noncu" ]; then
ncu --set full --clock-control none --import-source on -k regex:vox_cluster -s 3 -c 1 -f -o gpurun_out/prof_vox_$tag python tools/vox_probe.py 128 2 > gpurun_out/ncu_vox.log 2>&1
tail -1 gpurun_out/ncu_vox.log
fi
