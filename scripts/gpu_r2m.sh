#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_forward.py tests/test_gpu_bf16_parity.py tests/test_gpu_tc.py::test_attention_tc_pingpong -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -n 6
bash scripts/gpu_ab_bench.sh BD_LAST_LAYER_PRUNE=0 BD_LAST_LAYER_PRUNE=1 BD_LAST_LAYER_PRUNE=0 BD_LAST_LAYER_PRUNE=1
