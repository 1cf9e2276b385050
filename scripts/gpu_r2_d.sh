#!/bin/bash
# Round 2, call D (1 GPU): full GPU suite after the prefetch fix, prefetch decision trace, kernel micro-benchmarks (spmv 16 vs 8
# consumer warps), C2 bench (stock + fused), racecheck on a minimal textbook TMA pipeline, ncu --set full of both spmv variants.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r2d_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r2d_pytest_gpu.log | cut -c1-300
TB_PF_DEBUG=1 timeout 120 python scripts/pf_debug.py 2> gpurun_out/r2d_pf_debug.log > /dev/null; echo "pf_debug rc=$?"; grep -c served gpurun_out/r2d_pf_debug.log; grep -n "drop\|bad" gpurun_out/r2d_pf_debug.log | head
timeout 600 python scripts/bench_kernels.py > gpurun_out/r2d_bench_kernels.json 2> gpurun_out/r2d_bench_kernels.err; echo "bench_kernels rc=$?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/r2d_bench_kernels.json"))
for r in d["rows"]:
    print("%-100s %8.4f ms %8.1f GB/s %.3f" % (r["kernel"][:100], r["ms"], r["gbs"], r["frac_of_measured_hbm_peak"]))
PY
tail -3 gpurun_out/r2d_bench_kernels.err
for route in fused stock; do
  timeout 600 python bench.py --workload c2_qp_n8192_m8192_p1024 --route $route --steps 200 --no-cpu-baseline --no-parity > gpurun_out/r2d_bench_c2_$route.json 2> gpurun_out/r2d_bench_c2_$route.err; echo "bench c2 $route rc=$?: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2d_bench_c2_$route.json) $(grep -o '"e2e": {"value": [0-9.]*' gpurun_out/r2d_bench_c2_$route.json)"; tail -2 gpurun_out/r2d_bench_c2_$route.err
done
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -o /tmp/rc_min scripts/racecheck_tma_minimal.cu && /tmp/rc_min
timeout -k 10 200 compute-sanitizer --tool racecheck /tmp/rc_min > gpurun_out/r2d_racecheck_minimal.log 2>&1; echo "racecheck minimal rc=$?"; head -c 3000 gpurun_out/r2d_racecheck_minimal.log
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:'spmv_stream_kernel' -c 4 -f -o gpurun_out/r2d_spmv python scripts/sp_only.py once > gpurun_out/r2d_ncu_spmv.out 2>&1; echo "ncu spmv rc=$?"; tail -2 gpurun_out/r2d_ncu_spmv.out
du -sh gpurun_out
