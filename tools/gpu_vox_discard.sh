#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q -k "voxel" 2>&1 | tail -3
for lib in "" nodiscard; do
for cf in 32 48 64; do
L=""; [ -n "$lib" ] && L="D3D_B200_LIB=$PWD/d3d_b200/libd3d_b200_$lib.so"
env $L D3D_B200_VOX_CF=$cf python tools/vox_probe.py 128 20 2>&1 | tail -1 | sed "s/^/cf=$cf /" | cut -c1-40,100-200
done
done
for cf in 32 64; do
D3D_B200_VOX_CF=$cf ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:vt_ -s 24 -c 16 --csv --log-file gpurun_out/vt_dram_$cf.csv python tools/vox_probe.py 128 3 > /dev/null 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/vt_dram_$cf.csv')) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value')
agg=collections.defaultdict(lambda: collections.defaultdict(float)); n=collections.Counter()
for r in rows[1:]:
    agg[r[ki][:20]][r[mi]] += float(r[vi].replace(',','')); n[r[ki][:20]]+=1
tot=0
for k,v in agg.items():
    print('cf=$cf', k, n[k]//3, {m: round(x/1e6,1) for m,x in v.items()}); tot+=v['dram__bytes_read.sum']+v['dram__bytes_write.sum']
print('cf=$cf total MB over', sum(n.values())//3, 'launches', round(tot/1e6,1))
PY
done
