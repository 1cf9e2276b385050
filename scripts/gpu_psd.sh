#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_cone_eig_gpu.py tests/test_level1_gpu.py -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/test_cone.log
echo "== cone tests exit ${PIPESTATUS[0]}"; tail -12 gpurun_out/test_cone.log
timeout -k 10 600 python scripts/bench_psd.py 512 20 > gpurun_out/bench_psd.json 2> gpurun_out/bench_psd.err
echo "== psd bench exit $?"; cat gpurun_out/bench_psd.json; tail -3 gpurun_out/bench_psd.err
timeout -k 10 600 python bench.py --steps 200 --warmup 10 --workload socp_small_128x64_A8192x4096 --no-cpu-baseline > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err
echo "== bench small exit $?"; head -c 1500 gpurun_out/bench_small.json; tail -3 gpurun_out/bench_small.err
