#!/bin/bash
# BASELINE configs 4 and 5 through bench.py (driver-reproducible), plus the dense / cached side benches
mkdir -p gpurun_out
echo "=== config 4"; timeout 900 python bench.py --config 4 --steps 2 --warmup 1 > gpurun_out/bench_config4.json 2> gpurun_out/bench_config4.err; echo "exit $?"; cut -c1-600 gpurun_out/bench_config4.json; tail -n 3 gpurun_out/bench_config4.err
echo "=== config 5"; timeout 900 python bench.py --config 5 --steps 3 --warmup 1 > gpurun_out/bench_config5.json 2> gpurun_out/bench_config5.err; echo "exit $?"; cut -c1-900 gpurun_out/bench_config5.json; tail -n 3 gpurun_out/bench_config5.err
echo "=== dense"; timeout 600 python scripts/bench_dense.py > gpurun_out/bench_dense.json 2>&1; tail -n 2 gpurun_out/bench_dense.json | cut -c1-400
echo "=== cached"; timeout 600 python scripts/bench_cached.py > gpurun_out/bench_cached.json 2>&1; tail -n 2 gpurun_out/bench_cached.json | cut -c1-400
