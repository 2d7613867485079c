#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q -k "nms" 2>&1 | tail -5
for path in dense edges; do
D3D_B200_NMS_BATCH_PATH=$path timeout 300 python bench.py --op c5 --no-cpu-baseline --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$path c5 ms', d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'])"
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c5.csv timeout 300 python bench.py --op c5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c5_under_ncu.log 2>&1
python profiles/summarize_launches.py gpurun_out/launches_c5.csv | head -20
