#!/bin/bash
# tuning helper (run under gpurun): voxel parity tests, C2 probe timing per back end, launch list of the tile pipeline
tag=${1:-x}
timeout 900 python -m pytest tests -m gpu -x -q -k "voxel" 2>&1 | tail -5
for cf in 4 8 16; do
D3D_B200_VOX_CF=$cf python tools/vox_probe.py 128 20 2>&1 | tail -1 | sed "s/^/cf=$cf /" | tee -a gpurun_out/vox_$tag.txt
done
python tools/vox_probe.py 1 20 2>&1 | tail -1 | tee -a gpurun_out/vox_$tag.txt
python tools/vox_probe.py 128 20 auto_no_tiles 2>&1 | tail -1 | sed "s/^/cluster /" | tee -a gpurun_out/vox_$tag.txt
python tools/vox_probe.py 1 20 auto_no_tiles 2>&1 | tail -1 | sed "s/^/cluster /" | tee -a gpurun_out/vox_$tag.txt
if [ "$2" != "noncu" ]; then
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:vt_ -s 60 -c 20 --csv --log-file gpurun_out/vox_launches_$tag.csv python tools/vox_probe.py 128 2 > gpurun_out/ncu_vox.log 2>&1
tail -2 gpurun_out/ncu_vox.log
fi
