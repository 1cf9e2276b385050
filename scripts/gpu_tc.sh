#!/bin/bash
# First light for the tcgen05 GEMM: each case in its own process under a timeout.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
for args in "128 1" "512 1" "132 1" "64 1" "512 2" "512 4" "320 4" "1024 0" "512 1 1"; do
  timeout 120 python scripts/tc_check.py $args || echo "FAILED/timeout rc=$? : $args"
done
echo "--- descriptor swap variant"
TB_TC_DESC_SWAP=1 timeout 120 python scripts/tc_check.py 128 1 || echo "swap variant failed rc=$?"
echo "--- psd projection timing"
timeout 300 python scripts/bench_psd.py 512 20
echo "--- pytest cone/eig"
timeout 900 python -m pytest tests/test_cone_eig_gpu.py -x -q -m gpu 2>&1 | tail -15
} > gpurun_out/tc_first.log 2>&1
tail -40 gpurun_out/tc_first.log
