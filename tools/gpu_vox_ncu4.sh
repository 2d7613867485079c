#!/bin/bash
tag=${1:-x}
D3D_B200_VOX_CF=128 ncu --set full --clock-control none --cache-control none --import-source on -k regex:vt_ -s 12 -c 4 -f -o gpurun_out/prof_vt4_$tag python tools/vox_probe.py 128 2 > gpurun_out/ncu_vt4.log 2>&1
tail -2 gpurun_out/ncu_vt4.log
