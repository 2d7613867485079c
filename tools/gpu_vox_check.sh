#!/bin/bash
# tuning helper (run under gpurun): voxel parity tests, C2 probe timing per back end, launch list of the tile pipeline
tag=${1:-x}
timeout 900 python -m pytest tests -m gpu -x -q -k "voxel" 2>&1 | tail -5
for cf in 4 8 16; do
D3D_B200_VOX_CF=$cf python tools/vox_probe.py 128 20 2>&1 | tail -1 | sed "s/^/cf=$cf /" | tee -a gpurun_out/vox_$tag.txt
done
python tools/vox_probe.py 1 20 2>&1 | tail -1 | tee -a gpurun_out/vox_$tag.txt
if [ "$2" != "noncu" ]; then
ncu --set full --clock-control none --import-source on -k regex:vt_tick -s 66 -c 1 -f -o gpurun_out/prof_vt_$tag python tools/vox_probe.py 128 2 > gpurun_out/ncu_vt.log 2>&1
tail -2 gpurun_out/ncu_vt.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none -k regex:vt_ -s 60 -c 20 --csv --log-file gpurun_out/vox_launches_$tag.csv python tools/vox_probe.py 128 2 > gpurun_out/ncu_vox.log 2>&1
fi
