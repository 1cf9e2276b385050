#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_solver_gpu.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/test_solver_gpu.log
echo "== solver tests exit ${PIPESTATUS[0]}"; tail -5 gpurun_out/test_solver_gpu.log
BENCH_DEBUG=1 timeout -k 10 900 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
echo "== bench c3 exit $?"; cat gpurun_out/bench_c3.json | head -c 3000; tail -8 gpurun_out/bench_c3.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
bash scripts/gpu_prof.sh
