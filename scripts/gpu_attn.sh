#!/bin/bash
# attention iteration loop: parity tests of the ping-pong kernel, kernel micro-bench, role timeline
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py::test_attention_tc_pingpong -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -n 5
timeout 300 python scripts/bench_kernels.py attn > gpurun_out/bench_kernels_attn.json 2>&1; cat gpurun_out/bench_kernels_attn.json
timeout 120 python scripts/trace_attention.py > gpurun_out/trace_attn.txt 2>&1; sed -n 1,30p gpurun_out/trace_attn.txt
timeout 120 python scripts/trace_attention.py 384 12 64 261 > gpurun_out/trace_attn_dino.txt 2>&1; sed -n 1,14p gpurun_out/trace_attn_dino.txt
