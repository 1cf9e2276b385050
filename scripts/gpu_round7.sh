#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
for w in c3_socp_1024x64_A65536x16384 c2_qp_n8192_m8192_p1024 c4_sdp_psd512_A131328x1024; do
  timeout 600 python bench.py --workload $w --steps 200 --no-cpu-baseline > gpurun_out/bench_${w}_wide.json 2> gpurun_out/bench_${w}_wide.err
  echo "$w rc=$?: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_${w}_wide.json) $(grep -o '"gpu_launches": [0-9]*' gpurun_out/bench_${w}_wide.json)"
  tail -2 gpurun_out/bench_${w}_wide.err
done
