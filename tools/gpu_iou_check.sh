#!/bin/bash
# tuning helper (run under gpurun): IoU parity tests, C4 bench line, one ncu --set full capture of the tile kernel
tag=${1:-x}
timeout 300 python -m pytest tests -m gpu -x -q -k "iou" 2>&1 | tail -3
python bench.py --op iou --no-cpu-baseline --steps 5 2>&1 | tail -1 > gpurun_out/iou_$tag.json
python -c "
import json; d=json.load(open('gpurun_out/iou_$tag.json')); o=d['ops']['iou'] if 'ops' in d else d; print(o['ms_per_step'], o['roofline']['achieved'], o['roofline']['peak'], o['roofline']['frac'])"
ncu --set full --clock-control none --import-source on -k regex:iou2dr_tile -s 2 -c 1 -f -o gpurun_out/prof_iou_$tag python bench.py --op iou --no-cpu-baseline --steps 2 --warmup 1 --iou-n 30000 > gpurun_out/ncu_iou.log 2>&1
tail -1 gpurun_out/ncu_iou.log
