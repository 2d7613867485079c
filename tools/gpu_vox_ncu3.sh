#!/bin/bash
tag=${1:-x}
D3D_B200_VOX_CF=8 ncu --set full --clock-control none --cache-control none --import-source on -k regex:vt_tick -s 76 -c 2 -f -o gpurun_out/prof_vt3_$tag python tools/vox_probe.py 128 2 > gpurun_out/ncu_vt3.log 2>&1
tail -2 gpurun_out/ncu_vt3.log
