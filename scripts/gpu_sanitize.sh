#!/bin/bash
# compute-sanitizer passes over small-shape tests (memcheck: out-of-bounds / misaligned; racecheck: shared-memory hazards):
# the attention kernel, top-20, the warp-per-query PnP, and one bf16 forward that runs the query-window last block and the graphs.
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  echo "=== $tool"
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python -m pytest "tests/test_gpu_tc.py::test_attention_tc_pingpong" "tests/test_gpu_simt.py::test_corners_topk_ties_and_degenerate_maps" "tests/test_gpu_simt.py::test_pnp_matches_cv2_fixture" -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/sanitize_$tool.log 2>&1
  echo "exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error|hazard" gpurun_out/sanitize_$tool.log | head -12
done
echo "=== memcheck: forward (query-window last block, packed record, graph replay)"
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest "tests/test_gpu_forward.py::test_last_decoder_block_on_query_rows_is_bit_identical" "tests/test_gpu_forward.py::test_forward_packed_record_equals_packed_forward" -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/sanitize_memcheck_forward.log 2>&1
echo "exit $?"; grep -E "ERROR SUMMARY|passed|failed|Error" gpurun_out/sanitize_memcheck_forward.log | head -12
