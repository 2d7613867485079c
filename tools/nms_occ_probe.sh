for v in "" pc3 pc4; do L=""; [ -n "$v" ] && L=$PWD/d3d_b200/libd3d_b200_$v.so
  D3D_B200_LIB=$L python bench.py --op nms --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('variant=$v', d['ms_per_step'], d['config']['kept'])"
done
