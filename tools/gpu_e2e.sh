#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q -k "voxel" 2>&1 | tail -3
python bench.py --op voxel --no-cpu-baseline --steps 10 > gpurun_out/bench_vox.json 2> gpurun_out/bench_vox.err; tail -2 gpurun_out/bench_vox.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_vox.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['frac'], d['e2e'], d['single_frame'], d['numa'])"
