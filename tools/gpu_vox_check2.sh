#!/bin/bash
tag=${1:-x}
timeout 900 python -m pytest tests -m gpu -x -q -k "voxel" 2>&1 | tail -5
ncu --set full --clock-control none --import-source on -k regex:vt_tick -s 66 -c 1 -f -o gpurun_out/prof_vt_$tag python tools/vox_probe.py 128 2 > gpurun_out/ncu_vt.log 2>&1
tail -2 gpurun_out/ncu_vt.log
