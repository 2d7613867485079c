#!/bin/bash
timeout 300 python -m pytest tests -m gpu -x -q -k "scatter" 2>&1 | tail -5
for cfg in "gather 1" "tiles 1"; do set -- $cfg
D3D_B200_SCATTER_PATH=$1 D3D_B200_SCATTER_PIPE=$2 timeout 300 python bench.py --op scatter --no-cpu-baseline --steps 10 > gpurun_out/bench_scatter.json 2> gpurun_out/bench_scatter.err; tail -2 gpurun_out/bench_scatter.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_scatter.json').read().strip().splitlines()[-1]); print('$1 pipe=$2', 'fwd+bwd ms', d['ms_per_step'], 'fwd ms', d['config']['forward_ms'], 'frac', d['roofline']['frac'], 'anchor', d['config']['forward_sum_anchor'])"
done
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_scatter.csv timeout 300 python bench.py --op scatter --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/scatter_under_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_scatter.csv')) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); ii=h.index('ID')
agg=collections.defaultdict(lambda: collections.defaultdict(list))
for r in rows[1:]:
    agg[r[ki][:60]][r[mi]].append(float(r[vi].replace(',','')))
for k,v in agg.items():
    if 'd3d' in k: print(k, {m:(len(x), round(sum(x)/len(x),1)) for m,x in v.items()})
PY
