#!/bin/bash
# round-end evidence: parity tests, smoke, default bench line, launch list of the bench command, ncu captures of the fp64 IoU kernel
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -3 gpurun_out/bench_full.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_full.json').read().strip().splitlines()[-1])
print({k:(v if not isinstance(v,(dict,list)) else '...') for k,v in d.items()})
print('roofline', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'single', d.get('single_frame'))
for op,r in d.get('ops',{}).items(): print(op, r['value'], r['ms_per_step'], r['roofline'].get('frac'), r['e2e']['value'], r.get('ref_cuda'), r.get('soft'))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -1 gpurun_out/launches_bench.csv | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:iou2dr_tile_kernel -s 1 -c 1 -f -o gpurun_out/prof_iou_f64 python bench.py --op iou_f64 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_iou_f64.log 2>&1
tail -2 gpurun_out/ncu_iou_f64.log
ncu --set full --clock-control none --import-source on -k regex:iou2dr_tile_kernel -s 1 -c 1 -f -o gpurun_out/prof_iou_f32 python bench.py --op iou --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_iou_f32.log 2>&1
tail -2 gpurun_out/ncu_iou_f32.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:nms -c 200 --csv --log-file gpurun_out/launches_c5.csv python bench.py --op c5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c5_under_ncu.log 2>&1
tail -1 gpurun_out/launches_c5.csv | cut -c1-200
