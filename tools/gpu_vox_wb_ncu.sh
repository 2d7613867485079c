#!/bin/bash
# tuning helper (run under gpurun): per-kernel times of the voxel tile pipeline for library variants (ncu launch list, cold, serialised)
for v in "" "$@"; do
  lib=$PWD/d3d_b200/libd3d_b200${v:+_$v}.so
  D3D_B200_LIB=$lib ncu --metrics gpu__time_duration.sum --clock-control none -k regex:vt_ -s 16 -c 16 --csv --log-file gpurun_out/vox_l_${v:-default}.csv python tools/vox_probe.py 128 2 > gpurun_out/ncu_vox_${v:-default}.log 2>&1
  python profiles/summarize_launches.py gpurun_out/vox_l_${v:-default}.csv | tee gpurun_out/vox_l_${v:-default}.txt
done
