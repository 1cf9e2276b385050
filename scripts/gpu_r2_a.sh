#!/bin/bash
# Round 2, call A: full GPU test suite (new C1 / full-size C3 / shim-protocol tests), the new bench harness (repo arm with
# parity + cpu_baseline, reference arm at full size), same-box cuBLAS / cuSOLVER comparison, C4 / C2 parity legs.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc; free -g | head -2
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2a_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2a_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_c3.json 2> gpurun_out/r2a_bench_c3.err; echo "bench c3 rc=$?"; cut -c1-400 gpurun_out/r2a_bench_c3.json; tail -3 gpurun_out/r2a_bench_c3.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err; echo "bench ref rc=$?"; cut -c1-300 gpurun_out/r2a_bench_ref.json; tail -3 gpurun_out/r2a_bench_ref.err
timeout 600 python scripts/bench_vs_cuda_libs.py > gpurun_out/r2a_vs_cuda_libs.json 2> gpurun_out/r2a_vs_cuda_libs.err; echo "vs libs rc=$?"; cut -c1-600 gpurun_out/r2a_vs_cuda_libs.json; tail -3 gpurun_out/r2a_vs_cuda_libs.err
timeout 600 python bench.py --workload c4_sdp_psd512_A131328x1024 --steps 100 --parity-k 100 --no-cpu-baseline > gpurun_out/r2a_bench_c4.json 2> gpurun_out/r2a_bench_c4.err; echo "bench c4 rc=$?"; grep -o '"parity": {.*' gpurun_out/r2a_bench_c4.json | cut -c1-900; tail -3 gpurun_out/r2a_bench_c4.err
timeout 600 python bench.py --workload c2_qp_n8192_m8192_p1024 --steps 100 --parity-k 100 --no-cpu-baseline > gpurun_out/r2a_bench_c2.json 2> gpurun_out/r2a_bench_c2.err; echo "bench c2 rc=$?"; grep -o '"parity": {.*' gpurun_out/r2a_bench_c2.json | cut -c1-900; tail -3 gpurun_out/r2a_bench_c2.err
timeout 600 python bench.py --steps 100 --shim-protocol 1 --no-cpu-baseline --no-parity > gpurun_out/r2a_bench_c3_shim.json 2> gpurun_out/r2a_bench_c3_shim.err; echo "bench c3 shim rc=$?"; cut -c1-200 gpurun_out/r2a_bench_c3_shim.json; tail -3 gpurun_out/r2a_bench_c3_shim.err
