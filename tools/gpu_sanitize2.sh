#!/bin/bash
# compute-sanitizer over the kernels changed in the last session of round 2: batched NMS (edge path), cooperative radix sort, IoU tile kernel
run() { tool=$1; shift; sel=$1; shift
  out=$(timeout 1200 compute-sanitizer --tool $tool "$@" python -m pytest tests -m gpu -x -q -k "$sel" 2>&1)
  echo "$tool | $sel | $(echo "$out" | grep -E "passed|failed|error" | tail -1) | $(echo "$out" | grep -E "ERROR SUMMARY|RACECHECK SUMMARY" | tail -1)"
  echo "$out" | grep -E "Race reported|Invalid|Uninitialized|hazard" | sort | uniq -c | head -8
}
{
run memcheck "nms_batch or sort_forms or nms_known or iou_c1 or iou_degenerate or box3d"
run racecheck "nms_batch or sort_forms or nms_known" --racecheck-report analysis
} 2>&1 | tee gpurun_out/sanitize_r2b.txt
