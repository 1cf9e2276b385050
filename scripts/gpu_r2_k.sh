#!/bin/bash
# Round 2, call K (1 GPU): full GPU suite (device block pool, timeline), real device timelines of C3 / C2 / C4, bench lines.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -q -m gpu > gpurun_out/r2k_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2k_pytest_gpu.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --timeline gpurun_out/r2k_timeline_c3.txt > gpurun_out/r2k_bench_c3_default.json 2> gpurun_out/r2k_bench_c3_default.err; echo "bench c3 default rc=$?"; cut -c1-200 gpurun_out/r2k_bench_c3_default.json; tail -2 gpurun_out/r2k_bench_c3_default.err
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-parity > gpurun_out/r2k_bench_c3_100.json 2> gpurun_out/r2k_bench_c3_100.err; echo "bench c3 100 rc=$?: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2k_bench_c3_100.json) $(grep -o '"e2e": {"value": [0-9.]*' gpurun_out/r2k_bench_c3_100.json)"
for route in fused stock; do
  timeout 600 python bench.py --workload c2_qp_n8192_m8192_p1024 --route $route --steps 200 --no-cpu-baseline --no-parity --timeline gpurun_out/r2k_timeline_c2_$route.txt > gpurun_out/r2k_bench_c2_$route.json 2> gpurun_out/r2k_bench_c2_$route.err; echo "bench c2 $route rc=$?: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2k_bench_c2_$route.json) $(grep -o '"e2e": {"value": [0-9.]*' gpurun_out/r2k_bench_c2_$route.json)"; tail -2 gpurun_out/r2k_bench_c2_$route.err
done
timeout 600 python bench.py --workload c4_sdp_psd512_A131328x1024 --steps 100 --no-cpu-baseline --no-parity --timeline gpurun_out/r2k_timeline_c4.txt > gpurun_out/r2k_bench_c4.json 2> gpurun_out/r2k_bench_c4.err; echo "bench c4 rc=$?: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2k_bench_c4.json) $(grep -o '"e2e": {"value": [0-9.]*' gpurun_out/r2k_bench_c4.json)"
timeout -k 5 600 python scripts/bench_kernels.py > gpurun_out/r2k_bench_kernels.json 2> gpurun_out/r2k_bench_kernels.err; echo "bench_kernels rc=$?"; grep -o '"kernel": "transform_sp[^}]*' gpurun_out/r2k_bench_kernels.json | cut -c1-260
du -sh gpurun_out
