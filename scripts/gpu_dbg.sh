#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/tests_gpu_all.log
echo "== tests exit ${PIPESTATUS[0]}"; tail -30 gpurun_out/tests_gpu_all.log
for PF in 1 0; do
  timeout -k 10 600 python bench.py --steps 100 --warmup 10 --pair-fusion $PF --no-cpu-baseline > gpurun_out/bench_pf$PF.json 2> gpurun_out/bench_pf$PF.err
  echo "== bench pf=$PF exit $?"; head -c 3000 gpurun_out/bench_pf$PF.json; tail -3 gpurun_out/bench_pf$PF.err
done
timeout -k 10 600 python scripts/bench_psd.py 512 20 > gpurun_out/bench_psd.json 2> gpurun_out/bench_psd.err
echo "== psd bench exit $?"; cat gpurun_out/bench_psd.json; tail -3 gpurun_out/bench_psd.err
