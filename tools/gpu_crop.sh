#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q -k "crop" 2>&1 | tail -3
for path in 0; do
D3D_B200_CROP_PATH=$path python bench.py --op crop --no-cpu-baseline --steps 10 > gpurun_out/bench_crop.json 2> gpurun_out/bench_crop.err; tail -2 gpurun_out/bench_crop.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_crop.json').read().strip().splitlines()[-1]); print('path $path', d['ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step'])"
done
