#!/bin/bash
# One gpurun call: per-file GPU parity tests (each under its own timeout so a hung kernel cannot eat the box),
# then a short bench.  Logs land in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for f in test_level1_gpu test_gemv_gpu test_cone_eig_gpu test_solver_gpu; do
  timeout -k 10 ${TEST_TIMEOUT:-600} python -m pytest tests/$f.py -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -${TAIL:-60} > gpurun_out/$f.log
  echo "== $f: exit ${PIPESTATUS[0]}" | tee -a gpurun_out/summary.txt
  tail -3 gpurun_out/$f.log
done
if [ -n "$BENCH_ARGS" ]; then
  timeout -k 10 900 python bench.py $BENCH_ARGS > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "== bench exit $?" | tee -a gpurun_out/summary.txt
  cat gpurun_out/bench.json | head -c 3000
  tail -5 gpurun_out/bench.err
fi
