#!/bin/bash
# tuning helper (run under gpurun): NMS / sort / voxel parity tests, then the NMS bench with and without the cooperative sort
timeout 900 python -m pytest tests -m gpu -x -q -k "nms or sort or voxel" 2>&1 | tail -5
for coop in 0 1; do
D3D_B200_SORT_COOP=$coop timeout 300 python bench.py --op nms --no-cpu-baseline --steps 10 > gpurun_out/bench_nms.json 2> gpurun_out/bench_nms.err; tail -2 gpurun_out/bench_nms.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_nms.json').read().strip().splitlines()[-1]); r=d['roofline']; print('coop=$coop', 'ms', d['ms_per_step'], 'sort', r['ms_sort_gather'], 'cand', r['ms_candidates'], 'resolve', r['ms_resolve'], 'kept', d['config']['kept'], 'e2e', d['e2e']['ms_per_step'])"
done
