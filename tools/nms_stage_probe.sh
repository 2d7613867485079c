# tuning probe: NMS C3 step time against the resolve kernel's thread count and list staging depth
for nt in 64 128 256 512 1024; do
  D3D_B200_NMS_NT=$nt python bench.py --op nms --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('nt=$nt', d['ms_per_step'], d['config']['kept'])"
done
