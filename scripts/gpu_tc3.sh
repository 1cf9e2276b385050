#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for args in "128 1" "512 8" "132 2" "320 4"; do
  timeout 120 python scripts/tc_check.py $args || echo "FAILED/timeout rc=$? : $args"
done
TB_TC_PDL=0 timeout 200 python scripts/tc_trace.py 512
PSD_BENCH_JACOBI=0 timeout 300 python scripts/bench_psd.py 512 20
timeout 900 python -m pytest tests/test_cone_eig_gpu.py -x -q -m gpu 2>&1 | tail -3
} > gpurun_out/tc_third.log 2>&1
cat gpurun_out/tc_third.log
