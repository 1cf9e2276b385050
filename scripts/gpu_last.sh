#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemv_gpu.py tests/test_solver_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --steps 200 --no-cpu-baseline > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err; echo "c3: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_last.json) $(grep -o '"avg_launch_ms": [0-9.]*' gpurun_out/bench_last.json)"
