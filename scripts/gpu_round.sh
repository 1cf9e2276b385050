#!/bin/bash
# One gpurun call for a whole 1-GPU check: parity tests, smoke, bench (both arms), ncu launch list + full capture.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
  timeout -k 10 ${TEST_TIMEOUT:-1200} python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/tests_gpu.log
  echo "== tests exit ${PIPESTATUS[0]}"; tail -4 gpurun_out/tests_gpu.log
fi
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout -k 10 900 python bench.py ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "== bench exit $?"; head -c 4000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ -z "$SKIP_REF" ]; then
  timeout -k 10 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
  echo "== bench ref exit $?"; head -c 2000 gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
fi
if [ -z "$SKIP_PROF" ]; then bash scripts/gpu_prof.sh; fi
