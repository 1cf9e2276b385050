#!/bin/bash
# Round 2, call L (1 GPU): cost of a vector-program stage; A/B of the cluster / wide threshold (48 K vs 256 K elements) on C3, C2, C4
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python scripts/bench_vprog_stage.py > gpurun_out/r2l_vprog_stage.log 2>&1; echo "vprog stage rc=$?"; cut -c1-220 gpurun_out/r2l_vprog_stage.log | tail -22
for mx in 49152 262144; do
  TB_VPROG_MAX_N=$mx timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-parity --timeline gpurun_out/r2l_timeline_c3_$mx.txt > gpurun_out/r2l_bench_c3_$mx.json 2> gpurun_out/r2l_bench_c3_$mx.err; echo "bench c3 max_n=$mx rc=$?: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2l_bench_c3_$mx.json) $(grep -o '"gpu_launches": [0-9]*' gpurun_out/r2l_bench_c3_$mx.json)"; tail -1 gpurun_out/r2l_bench_c3_$mx.err | cut -c1-200
  TB_VPROG_MAX_N=$mx timeout 600 python bench.py --workload c2_qp_n8192_m8192_p1024 --steps 200 --no-cpu-baseline --no-parity > gpurun_out/r2l_bench_c2_$mx.json 2> gpurun_out/r2l_bench_c2_$mx.err; echo "bench c2 max_n=$mx rc=$?: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2l_bench_c2_$mx.json)"
  TB_VPROG_MAX_N=$mx timeout 600 python bench.py --workload c4_sdp_psd512_A131328x1024 --steps 100 --no-cpu-baseline --no-parity > gpurun_out/r2l_bench_c4_$mx.json 2> gpurun_out/r2l_bench_c4_$mx.err; echo "bench c4 max_n=$mx rc=$?: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2l_bench_c4_$mx.json)"
done
timeout -k 10 1200 python -m pytest tests -q -m gpu > gpurun_out/r2l_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2l_pytest_gpu.log | cut -c1-300
