#!/bin/bash
# compute-sanitizer passes over the small-shape kernel tests (memcheck: out-of-bounds / misaligned; racecheck: shared-memory hazards)
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  echo "=== $tool"
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest "tests/test_gpu_tc.py::test_attention_tc_pingpong" "tests/test_gpu_simt.py::test_corners_topk_ties_and_degenerate_maps" -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/sanitize_$tool.log 2>&1
  echo "exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error|hazard" gpurun_out/sanitize_$tool.log | head -12
done
