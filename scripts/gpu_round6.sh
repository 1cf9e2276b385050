#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
for w in c3_socp_1024x64_A65536x16384 c2_qp_n8192_m8192_p1024 c4_sdp_psd512_A131328x1024; do
  for sp in 1 0; do
  timeout 600 python bench.py --workload $w --steps 200 --no-cpu-baseline --speculation $sp > gpurun_out/bench_${w}_spec$sp.json 2> gpurun_out/bench_${w}_spec$sp.err
  echo "$w spec=$sp rc=$?: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_${w}_spec$sp.json) $(grep -o '"pairs_served_without_reading_A_per_iteration": [0-9.]*' gpurun_out/bench_${w}_spec$sp.json) $(grep -o '"avg_launch_ms": [0-9.]*' gpurun_out/bench_${w}_spec$sp.json)"
  tail -2 gpurun_out/bench_${w}_spec$sp.err
  done
done
