#!/bin/bash
# tcgen05 GEMM: correctness cases each in its own process under a timeout, then timing variants.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
for args in "128 1" "512 1" "132 2" "512 2" "512 4" "512 8" "320 8" "1024 0"; do
  timeout 120 python scripts/tc_check.py $args || echo "FAILED/timeout rc=$? : $args"
done
echo "--- psd projection timing (default: PDL on, split-K up to 8)"
PSD_BENCH_JACOBI=0 timeout 300 python scripts/bench_psd.py 512 20
echo "--- PDL off"
TB_TC_PDL=0 PSD_BENCH_JACOBI=0 timeout 300 python scripts/bench_psd.py 512 20
echo "--- split-K capped at 4"
TB_TC_MAX_SPLITK=4 PSD_BENCH_JACOBI=0 timeout 300 python scripts/bench_psd.py 512 20
echo "--- pytest cone/eig"
timeout 900 python -m pytest tests/test_cone_eig_gpu.py -x -q -m gpu 2>&1 | tail -5
} > gpurun_out/tc_second.log 2>&1
tail -40 gpurun_out/tc_second.log
