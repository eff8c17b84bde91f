#!/bin/bash
# A/B of an environment switch on the same box: bash scripts/gpu_ab_bench.sh VAR=0 VAR=1 ...
mkdir -p gpurun_out
for setting in "$@"; do
  echo "== $setting"
  env $setting timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('q/s', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'attn frac', round(d['roofline']['frac'],3), 'gemm frac', round(d['roofline_gemm']['frac'],3), 'mhz', d['clocks']['sm_mhz'])
print({k: round(v,2) for k,v in d['kernel_ms_per_step'].items()})"
done
