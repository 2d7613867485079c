#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q -k "iou or box3d" 2>&1 | tail -4
for op in iou iou_f64 dist3d; do
timeout 300 python bench.py --op $op --no-cpu-baseline --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$op ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],4))"
done
