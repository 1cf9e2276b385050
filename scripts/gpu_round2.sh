#!/bin/bash
# GPU tests + the three single-GPU bench workloads (C3 headline, C2 QP, C4 SDP).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench c3 rc=$?"; cut -c1-400 gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3.err
timeout 600 python bench.py --workload c2_qp_n8192_m8192_p1024 --steps 200 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"; cut -c1-400 gpurun_out/bench_c2.json; tail -3 gpurun_out/bench_c2.err
timeout 600 python bench.py --workload c4_sdp_psd512_A131328x1024 --steps 100 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench c4 rc=$?"; cut -c1-400 gpurun_out/bench_c4.json; tail -3 gpurun_out/bench_c4.err
