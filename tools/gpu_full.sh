#!/bin/bash
# full GPU check of a round: parity tests, smoke, default bench line
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -3 gpurun_out/bench_full.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_full.json').read().strip().splitlines()[-1])
print({k:(v if not isinstance(v,(dict,list)) else '...') for k,v in d.items()})
print('roofline', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'single', d.get('single_frame'))
for op,r in d.get('ops',{}).items(): print(op, r['value'], r['ms_per_step'], r['roofline'].get('frac'), r['e2e']['value'])
PY
