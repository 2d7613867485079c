#!/bin/bash
tag=${1:-x}
for cf in 16 32 64 128; do
D3D_B200_VOX_CF=$cf python tools/vox_probe.py 128 20 2>&1 | tail -1 | sed "s/^/cf=$cf /" | tee -a gpurun_out/vox_cf_$tag.txt
done
for r in 1 3 7 15; do
D3D_B200_VOX_CF=128 D3D_B200_VOX_ROLES=$r python tools/vox_probe.py 128 20 2>&1 | tail -1 | sed "s/^/cf=128 roles=$r /" | tee -a gpurun_out/vox_cf_$tag.txt
done
D3D_B200_VOX_CF=128 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --cache-control none --clock-control none -k regex:vt_ -s 12 -c 4 --csv --log-file gpurun_out/vox_launches_$tag.csv python tools/vox_probe.py 128 2 > gpurun_out/ncu_vox.log 2>&1
D3D_B200_VOX_CF=32 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --cache-control none --clock-control none -k regex:vt_ -s 48 -c 16 --csv --log-file gpurun_out/vox_launches32_$tag.csv python tools/vox_probe.py 128 2 > gpurun_out/ncu_vox.log 2>&1
