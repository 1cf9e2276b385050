#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
TB_TC_PDL=0 timeout 200 python scripts/tc_trace.py 512 2>&1 | head -3
timeout 600 python bench.py --workload c2_qp_n8192_m8192_p1024 --steps 200 > gpurun_out/bench_c2_fused.json 2> gpurun_out/bench_c2_fused.err; echo "bench c2 fused rc=$?"; cut -c1-300 gpurun_out/bench_c2_fused.json; tail -3 gpurun_out/bench_c2_fused.err
timeout 600 python bench.py --workload c2_qp_n8192_m8192_p1024 --route stock --steps 200 --no-cpu-baseline > gpurun_out/bench_c2_stock.json 2> gpurun_out/bench_c2_stock.err; echo "bench c2 stock rc=$?"; cut -c1-200 gpurun_out/bench_c2_stock.json; tail -3 gpurun_out/bench_c2_stock.err
