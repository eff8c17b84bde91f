#!/bin/bash
# A/B: kernel micro-bench of the attention variants built by scripts/build_variant.sh
mkdir -p gpurun_out
echo "== default"; timeout 300 python scripts/bench_kernels.py attn 2>&1 | grep -B1 tflops | grep -v "^--" | paste - - 
for l in scripts/_bin/lib_*.so; do echo "== $l"; BD_LIB_PATH=$l timeout 300 python scripts/bench_kernels.py attn 2>&1 | grep -B1 tflops | grep -v "^--" | paste - - ; done
