#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_level1_gpu.py tests/test_gemv_gpu.py -x -q -m gpu 2>&1 | tail -4
timeout 900 python -m pytest tests/test_solver_gpu.py tests/test_cone_eig_gpu.py -x -q -m gpu 2>&1 | tail -4
for w in socp_small_128x64_A8192x4096 c2_qp_n8192_m8192_p1024 c3_socp_1024x64_A65536x16384; do
  for v in 1 0; do
    timeout 600 python bench.py --workload $w --steps 200 --no-cpu-baseline --vprog $v > gpurun_out/bench_${w}_vprog$v.json 2> gpurun_out/bench_${w}_vprog$v.err
    echo "$w vprog=$v rc=$?: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_${w}_vprog$v.json) $(grep -o '"gpu_launches": [0-9]*' gpurun_out/bench_${w}_vprog$v.json) $(grep -o '"vector_programs": {[^}]*' gpurun_out/bench_${w}_vprog$v.json | cut -c1-120)"
    tail -2 gpurun_out/bench_${w}_vprog$v.err
  done
done
