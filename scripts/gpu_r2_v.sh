#!/bin/bash
# Round 2, call V (1 GPU): full suite with the long-finalize rule at its final threshold (512 K partial elements), C2 / small SOCP / C4 lines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2v_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2v_pytest_gpu.log | cut -c1-200
timeout 300 python bench.py --workload c2_qp_n8192_m8192_p1024 --steps 200 --no-cpu-baseline --parity-k 100 --timeline gpurun_out/r2v_timeline_c2.txt > gpurun_out/r2v_bench_c2.json 2> gpurun_out/r2v_bench_c2.err; echo "c2 rc=$?: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2v_bench_c2.json) $(grep -o '"e2e": {"value": [0-9.]*' gpurun_out/r2v_bench_c2.json) $(grep -o '"gpu_launches": [0-9]*' gpurun_out/r2v_bench_c2.json) $(grep -o '"parity": {"pass": [a-z]*' gpurun_out/r2v_bench_c2.json)"
timeout 300 python bench.py --workload socp_small_128x64_A8192x4096 --steps 200 --no-cpu-baseline --no-parity > gpurun_out/r2v_bench_small.json 2> gpurun_out/r2v_bench_small.err; echo "socp_small rc=$?: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2v_bench_small.json)"
timeout 300 python bench.py --workload c2_qp_n8192_m8192_p1024 --route stock --steps 200 --no-cpu-baseline --no-parity > gpurun_out/r2v_bench_c2_stock.json 2> gpurun_out/r2v_bench_c2_stock.err; echo "c2 stock rc=$?: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2v_bench_c2_stock.json)"
