#!/bin/bash
# Runs groups of -m gpu tests in separate processes (a trapped kernel poisons the CUDA context), each with a timeout,
# logging to gpurun_out/.  Usage (GPU box, repo root): bash scripts/gpu_ladder.sh [pytest node ids...]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/ladder_gpu.txt 2>&1
NODES="$@"
if [ -z "$NODES" ]; then
  NODES="tests/test_gpu_simt.py tests/test_gpu_tc.py::test_gemm_tc_f32 tests/test_gpu_tc.py::test_gemm_tc_epilogues tests/test_gpu_tc.py::test_qkv_project_tc tests/test_gpu_tc.py::test_attention_tc_pingpong tests/test_gpu_forward.py tests/test_gpu_bf16_parity.py tests/test_gpu_dense.py tests/test_metrics.py tests/test_checkpoint.py"
fi
i=0
for n in $NODES; do
  i=$((i+1))
  name=$(echo $n | sed 's#tests/##; s#\.py##; s#::#.#g')
  echo "=== $n"
  timeout 600 python -m pytest "$n" -m gpu -q --no-header -p no:cacheprovider -rA > gpurun_out/${i}_${name}.log 2>&1
  echo "exit $? ($n)"
  grep -E "PASSED|FAILED|ERROR|passed|failed|error" gpurun_out/${i}_${name}.log | tail -n 25
done
