#!/bin/bash
# Round 2, call F (1 GPU): transform_sp with the dynamic unit queue - parity, micro-benchmark (16 vs 8 warps), ncu, then the full suite.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gemv_gpu.py -q -m gpu -k "transform_sp" > gpurun_out/r2f_pytest_sp.log 2>&1; echo "pytest sp rc=$?"; tail -5 gpurun_out/r2f_pytest_sp.log | cut -c1-300
timeout -k 5 600 python scripts/bench_kernels.py > gpurun_out/r2f_bench_kernels.json 2> gpurun_out/r2f_bench_kernels.err; echo "bench_kernels rc=$?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/r2f_bench_kernels.json"))
for r in d["rows"]:
    print("%-100s %8.4f ms %8.1f GB/s %.3f" % (r["kernel"][:100], r["ms"], r["gbs"], r["frac_of_measured_hbm_peak"]))
PY
tail -3 gpurun_out/r2f_bench_kernels.err
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:'spmv_stream_kernel' -c 4 -f -o gpurun_out/r2f_spmv python scripts/sp_only.py once > gpurun_out/r2f_ncu_spmv.out 2>&1; echo "ncu spmv rc=$?"; tail -2 gpurun_out/r2f_ncu_spmv.out
timeout -k 10 1200 python -m pytest tests -q -m gpu > gpurun_out/r2f_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2f_pytest_gpu.log | cut -c1-300
for route in fused stock; do
  timeout 600 python bench.py --workload c2_qp_n8192_m8192_p1024 --route $route --steps 200 --no-cpu-baseline --no-parity > gpurun_out/r2f_bench_c2_$route.json 2> gpurun_out/r2f_bench_c2_$route.err; echo "bench c2 $route rc=$?: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2f_bench_c2_$route.json) $(grep -o '"e2e": {"value": [0-9.]*' gpurun_out/r2f_bench_c2_$route.json)"; tail -2 gpurun_out/r2f_bench_c2_$route.err
done
