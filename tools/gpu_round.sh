#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/vox_probe.py 128 20 2>&1 | tail -1
python bench.py --op iou --no-cpu-baseline --steps 5 > gpurun_out/bench_iou.json 2> gpurun_out/bench_iou.err; tail -2 gpurun_out/bench_iou.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_iou.json').read().strip().splitlines()[-1]); print(d.get('ref_cuda')); print(d['roofline'])"
python bench.py --op nms --no-cpu-baseline --steps 5 > gpurun_out/bench_nms.json 2> gpurun_out/bench_nms.err; tail -2 gpurun_out/bench_nms.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_nms.json').read().strip().splitlines()[-1]); print(d.get('ref_cuda')); print(d['roofline'])"
