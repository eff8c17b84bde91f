#!/bin/bash
# r01c probe: GELU formulation micro-benchmark + ncu full captures of the CTA-pair GEMM (DINO and decoder layers) and the DINO attention kernels.
mkdir -p gpurun_out
timeout 120 scripts/_bin/ubench_gelu > gpurun_out/ubench_gelu.txt 2>&1; cat gpurun_out/ubench_gelu.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel -s 30 -c 4 -o gpurun_out/prof_gemm2_dino -f python bench.py --steps 1 --warmup 3 --quick > gpurun_out/ncu_gemm2a.log 2>&1; echo "exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel -s 363 -c 4 -o gpurun_out/prof_gemm2_dec -f python bench.py --steps 1 --warmup 3 --quick > gpurun_out/ncu_gemm2b.log 2>&1; echo "exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_tc2_kernel<64>|attention_prefix_rows" -s 6 -c 2 -o gpurun_out/prof_attn_dino -f python bench.py --steps 1 --warmup 3 --quick > gpurun_out/ncu_attn_dino.log 2>&1; echo "exit $?"
ls -la gpurun_out/*.ncu-rep
