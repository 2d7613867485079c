#!/bin/bash
# compute-sanitizer over the kernels that are new in round 2 (parity tests at their small sizes)
run() { tool=$1; shift; sel=$1; shift
  out=$(timeout 1500 compute-sanitizer --tool $tool "$@" python -m pytest tests -m gpu -x -q -k "$sel" 2>&1)
  echo "$tool | $sel | $(echo "$out" | grep -E "passed|failed|error" | tail -1) | $(echo "$out" | grep -E "ERROR SUMMARY|RACECHECK SUMMARY" | tail -1)"
  echo "$out" | grep -E "Race reported|Invalid|Uninitialized|hazard" | sort | uniq -c | head -8
}
{
run memcheck "scatter_tile_path or box_crop or nms_soft or nms_batch or match_greedy or box_pdist or iou_differentiable"
run memcheck "voxel_golden_fixtures and auto and not no_tiles"
run racecheck "scatter_tile_path or box_crop or nms_soft or nms_batch or match_greedy or box_pdist or iou_differentiable" --racecheck-report analysis
run racecheck "voxel_golden_fixtures and auto and not no_tiles" --racecheck-report analysis
} 2>&1 | tee gpurun_out/sanitize_r2.txt
