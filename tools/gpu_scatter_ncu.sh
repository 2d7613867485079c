#!/bin/bash
ncu --set full --clock-control none --import-source on -k regex:st_tile_kernel -s 4 -c 2 -f -o gpurun_out/prof_st_tile python bench.py --op scatter --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_st_tile.log 2>&1
tail -2 gpurun_out/ncu_st_tile.log
