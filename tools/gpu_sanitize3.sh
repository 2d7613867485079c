#!/bin/bash
# compute-sanitizer memcheck over the kernels changed in the third session of round 2: NMS (cell-tile candidates, in-lists + pulled resolve, hi-half
# sort + run fix-up), batched NMS (Morton sort, buffered lists, pushed per-frame resolve), voxel bucket kernel, fp64 clip
run() { tool=$1; shift; sel=$1; shift
  out=$(timeout 900 compute-sanitizer --tool $tool "$@" python -m pytest tests -m gpu -x -q -k "$sel" 2>&1)
  echo "$tool | $sel | $(echo "$out" | grep -E " passed| failed|no tests ran" | tail -1) | $(echo "$out" | grep -E "ERROR SUMMARY|RACECHECK SUMMARY" | tail -1)"
  echo "$out" | grep -E "Race reported|Invalid|Uninitialized|hazard" | sort | uniq -c | head -8
}
{
run memcheck "nms_batch or nms_parallel or nms_back_ends or nms_known or sort_forms"
run memcheck "voxel_descending or voxel_c2 or iou_c1 or iou_degenerate"
} 2>&1 | tee gpurun_out/sanitize_r2c.txt
