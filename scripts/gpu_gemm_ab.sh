#!/bin/bash
# A/B: GEMM micro-bench of the library variants built by scripts/build_variant.sh
echo "== default"; timeout 300 python scripts/bench_kernels.py gemm 2>&1 | grep -v "^[{}]"
for l in scripts/_bin/lib_*.so; do echo "== $l"; BD_LIB_PATH=$l timeout 300 python scripts/bench_kernels.py gemm 2>&1 | grep -v "^[{}]"; done
