#!/bin/bash
# Round 2, call W (1 GPU): compute-sanitizer memcheck + synccheck over the tests of the kernels changed late in the round
# (transform_sp unit queue, one-launch cone pair, prefetch / reduction micro-ops, scalar-by-value operator)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SEL='not full_size and not c3 and not 5000 and not 512 and not 640 and not 2048 and not 70000 and not 16384'
timeout -k 10 400 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gemv_gpu.py tests/test_level1_gpu.py tests/test_cone_eig_gpu.py -q -x -k "$SEL" > gpurun_out/r2w_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r2w_sanitizer_memcheck.log
timeout -k 10 300 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_solver_gpu.py -q -x -k "scalar_prefetch or pair_fusion or vector_programs or speculative" > gpurun_out/r2w_sanitizer_memcheck_solver.log 2>&1; echo "memcheck solver rc=$?"; tail -3 gpurun_out/r2w_sanitizer_memcheck_solver.log
