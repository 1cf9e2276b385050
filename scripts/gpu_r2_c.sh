#!/bin/bash
# Round 2, call C: PDL / prefetch A-B, prefetch decision trace, sqrt, api trace, racecheck log, small ncu capture (outputs kept small).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x --deselect tests/test_solver_gpu.py::test_scalar_prefetch_halves_the_round_trips > gpurun_out/r2c_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r2c_pytest_gpu.log | cut -c1-300
timeout 300 python -m pytest tests/test_solver_gpu.py -q -m gpu -k scalar_prefetch > gpurun_out/r2c_pytest_prefetch.log 2>&1; echo "pytest prefetch rc=$?"; tail -5 gpurun_out/r2c_pytest_prefetch.log | cut -c1-300
TB_PF_DEBUG=1 timeout 120 python scripts/pf_debug.py 2> gpurun_out/r2c_pf_debug.log > /dev/null; echo "pf_debug rc=$?"; awk '/iteration 11/{p=1} p' gpurun_out/r2c_pf_debug.log | head -80
for cfg in "1 1" "0 1" "1 0" "0 0"; do
  set -- $cfg
  timeout 600 python bench.py --steps 100 --pdl $1 --scalar-prefetch $2 --no-cpu-baseline --no-parity > gpurun_out/r2c_bench_c3_pdl$1_pf$2.json 2> gpurun_out/r2c_bench_c3_pdl$1_pf$2.err; echo "bench c3 pdl=$1 prefetch=$2 rc=$?: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2c_bench_c3_pdl$1_pf$2.json) $(grep -o '"e2e": {"value": [0-9.]*' gpurun_out/r2c_bench_c3_pdl$1_pf$2.json)"; tail -2 gpurun_out/r2c_bench_c3_pdl$1_pf$2.err
  timeout 600 python bench.py --workload c2_qp_n8192_m8192_p1024 --steps 200 --pdl $1 --scalar-prefetch $2 --no-cpu-baseline --no-parity > gpurun_out/r2c_bench_c2_pdl$1_pf$2.json 2> gpurun_out/r2c_bench_c2_pdl$1_pf$2.err; echo "bench c2 pdl=$1 prefetch=$2 rc=$?: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2c_bench_c2_pdl$1_pf$2.json) $(grep -o '"e2e": {"value": [0-9.]*' gpurun_out/r2c_bench_c2_pdl$1_pf$2.json)"; tail -2 gpurun_out/r2c_bench_c2_pdl$1_pf$2.err
done
for pdl in 1 0; do
  timeout 600 python bench.py --workload c4_sdp_psd512_A131328x1024 --steps 100 --pdl $pdl --no-cpu-baseline --no-parity > gpurun_out/r2c_bench_c4_pdl$pdl.json 2> gpurun_out/r2c_bench_c4_pdl$pdl.err; echo "bench c4 pdl=$pdl rc=$?: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2c_bench_c4_pdl$pdl.json)"
done
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench_c3_default.json 2> gpurun_out/r2c_bench_c3_default.err; echo "bench c3 default rc=$?"; cut -c1-200 gpurun_out/r2c_bench_c3_default.json
timeout 600 python scripts/bench_sqrt.py > gpurun_out/r2c_bench_sqrt.json 2> gpurun_out/r2c_bench_sqrt.err; echo "bench_sqrt rc=$?"; cat gpurun_out/r2c_bench_sqrt.json | cut -c1-1500; tail -3 gpurun_out/r2c_bench_sqrt.err
timeout 300 python scripts/api_trace.py > gpurun_out/r2c_api_trace.json 2> gpurun_out/r2c_api_trace.err; echo "api_trace rc=$?"; tail -3 gpurun_out/r2c_api_trace.err
SEL='not full_size and not c3 and not 5000'
timeout -k 10 500 compute-sanitizer --tool racecheck --error-exitcode 99 python -m pytest tests/test_gemv_gpu.py -q -x -k "$SEL" > gpurun_out/r2c_racecheck_full.log 2>&1; echo "racecheck gemv rc=$?"; grep -v "^\.\|^$" gpurun_out/r2c_racecheck_full.log | head -c 6000 > gpurun_out/r2c_sanitizer_racecheck_gemv.log; rm -f gpurun_out/r2c_racecheck_full.log; head -60 gpurun_out/r2c_sanitizer_racecheck_gemv.log | cut -c1-250
timeout -k 10 300 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gemv_gpu.py tests/test_level1_gpu.py tests/test_cone_eig_gpu.py -q -x -k "$SEL and not 512 and not 640 and not 2048" > gpurun_out/r2c_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r2c_sanitizer_memcheck.log
timeout -k 10 300 compute-sanitizer --tool synccheck --error-exitcode 99 python -m pytest tests/test_gemv_gpu.py tests/test_cone_eig_gpu.py -q -x -k "$SEL and not 512 and not 640 and not 2048" > gpurun_out/r2c_sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -3 gpurun_out/r2c_sanitizer_synccheck.log
timeout -k 10 600 ncu --set full --clock-control none -k regex:'reduce_kernel|finalize|gemv_._generic|axpby_kernel|diag_kernel|cone_kernel|prefetch_reduce|vprog|stream_kernel.*Lb1' -c 30 -f -o gpurun_out/r2c_small_kernels python scripts/ncu_targets.py > gpurun_out/r2c_ncu_small.out 2>&1; echo "ncu small kernels rc=$?"; tail -2 gpurun_out/r2c_ncu_small.out; ls -la gpurun_out/r2c_small_kernels.ncu-rep
du -sh gpurun_out
