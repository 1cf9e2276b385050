#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_solver_gpu.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -80 > gpurun_out/test_solver_gpu.log
echo "== solver tests exit ${PIPESTATUS[0]}"; tail -30 gpurun_out/test_solver_gpu.log
BENCH_DEBUG=1 timeout -k 10 600 python bench.py --steps 20 --warmup 8 --workload socp_small_128x64_A8192x4096 --no-cpu-baseline > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err
echo "== bench small exit $?"; cat gpurun_out/bench_small.json | head -c 1500; tail -12 gpurun_out/bench_small.err
BENCH_DEBUG=1 timeout -k 10 900 python bench.py --steps 50 --warmup 8 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
echo "== bench c3 exit $?"; cat gpurun_out/bench_c3.json | head -c 3000; tail -12 gpurun_out/bench_c3.err
