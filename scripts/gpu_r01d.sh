#!/bin/bash
# r01d re-baseline: GPU tests, smoke, bench (both arms), kernel micro-bench, attention role timeline (no ncu).
mkdir -p gpurun_out
bash scripts/gpu_ladder.sh
echo "=== smoke"; timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "exit $?"; tail -n 3 gpurun_out/smoke.log
echo "=== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "exit $?"; cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
echo "=== kernels"; timeout 600 python scripts/bench_kernels.py > gpurun_out/bench_kernels.json 2>&1; cat gpurun_out/bench_kernels.json
echo "=== trace"; timeout 300 python scripts/trace_attention.py > gpurun_out/trace_attn.txt 2>&1; head -40 gpurun_out/trace_attn.txt
echo "=== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "exit $?"; cat gpurun_out/bench_ref.json
