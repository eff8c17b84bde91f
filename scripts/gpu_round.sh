#!/bin/bash
# One GPU session: tests, smoke, bench (both arms), kernel micro-bench, ncu launch list + full captures of the hot kernels.
mkdir -p gpurun_out
bash scripts/gpu_ladder.sh "$@"
echo "=== smoke"; timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "exit $?"; tail -n 3 gpurun_out/smoke.log
echo "=== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "exit $?"; cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
echo "=== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "exit $?"; cat gpurun_out/bench_ref.json
echo "=== kernels"; timeout 600 python scripts/bench_kernels.py > gpurun_out/bench_kernels.json 2>&1; echo "exit $?"
if [ "$NO_NCU" != "1" ]; then
echo "=== ncu launch list (219 launches per step; ~600 set-up launches + 3 warm-up steps skipped)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1260 -c 440 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --quick > gpurun_out/ncu_launch.log 2>&1; echo "exit $?"
echo "=== ncu full: decoder attention"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_tc2_kernel -s 60 -c 2 -o gpurun_out/prof_attn -f python bench.py --steps 1 --warmup 3 --quick > gpurun_out/ncu_attn.log 2>&1; echo "exit $?"
echo "=== ncu full: DINOv2 attention (main + prefix rows)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attn_tc2_kernel<64>|attention_prefix_rows" -s 6 -c 2 -o gpurun_out/prof_attn_dino -f python bench.py --steps 1 --warmup 3 --quick > gpurun_out/ncu_attn_dino.log 2>&1; echo "exit $?"
echo "=== ncu full: CTA-pair GEMM, decoder layer (qkv, proj, fc1, fc2)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel -s 363 -c 5 -o gpurun_out/prof_gemm -f python bench.py --steps 1 --warmup 3 --quick > gpurun_out/ncu_gemm.log 2>&1; echo "exit $?"
fi
ls -la gpurun_out | head -40
