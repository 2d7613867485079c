#!/bin/bash
# bench at N GPUs of one box: bash tools/gpu_scale.sh N
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -3 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
print("voxel", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"].get("ms_per_step"), d.get("numa"))
for op,r in d.get("ops",{}).items(): print(op, r["value"], r["ms_per_step"], r["e2e"]["value"], r.get("gather"))
PY
