#!/bin/bash
# A/B of attention-kernel build variants on one box: parity test + isolated timing (+ step time with STEP=1) per library.
#   bash scripts/gpu_attn_variants.sh default poly4 poly3 ...      (names of scripts/_bin/lib_<name>.so; "default" = in-tree library)
mkdir -p gpurun_out
for name in "$@"; do
  lib=""; [ "$name" != "default" ] && lib=scripts/_bin/lib_${name}.so
  echo "=== $name"
  BD_LIB_PATH=$lib timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q --no-header -p no:cacheprovider -x -k "attention" 2>&1 | tail -n 2
  BD_LIB_PATH=$lib timeout 300 python scripts/bench_kernels.py attn 2>&1 | python -c "
import sys, json
d = json.load(sys.stdin)
print(' '.join(f\"{k.split('_')[1]}:{v['v2']['ms']:.4f}ms/{v['v2']['tflops']:.0f}TF\" for k, v in d.items()))"
  if [ "$STEP" == "1" ]; then
    BD_LIB_PATH=$lib timeout 600 python bench.py --steps 10 --warmup 3 --quick 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('step: q/s', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'attn frac', round(d['roofline']['frac'],3), 'attn ms', round(d['kernel_ms_per_step']['attention'],2), 'dino attn ms', round(d['kernel_ms_per_step']['attention_dino'],2), 'mhz', d['clocks']['sm_mhz'])"
  fi
done
