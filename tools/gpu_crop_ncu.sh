#!/bin/bash
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_crop.csv python bench.py --op crop --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/crop_under_ncu.log 2>&1
python profiles/summarize_launches.py gpurun_out/launches_crop.csv | head -30
ncu --set full --clock-control none --import-source on -k regex:crop_hits_kernel -s 2 -c 1 -f -o gpurun_out/prof_crop_rows python bench.py --op crop --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_crop_rows.log 2>&1
tail -2 gpurun_out/ncu_crop_rows.log
