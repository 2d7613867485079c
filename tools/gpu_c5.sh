#!/bin/bash
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c5.csv timeout 300 python bench.py --op c5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c5_under_ncu.log 2>&1
python profiles/summarize_launches.py gpurun_out/launches_c5.csv | head -24
timeout 300 python bench.py --op c5 --no-cpu-baseline --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c5 ms', d['ms_per_step'], d['value'], d['e2e'])"
