#!/bin/bash
# Round 2, call G (1 GPU): full GPU suite; transform_sp (unit queue, 8 units per CTA, PDL) micro-benchmark + ncu; launch lists
# of the default C3 bench, C2 fused / stock and C4; bench lines of C2 / C4 / C3.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -q -m gpu > gpurun_out/r2g_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2g_pytest_gpu.log | cut -c1-300
timeout -k 5 600 python scripts/bench_kernels.py > gpurun_out/r2g_bench_kernels.json 2> gpurun_out/r2g_bench_kernels.err; echo "bench_kernels rc=$?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/r2g_bench_kernels.json"))
for r in d["rows"]:
    print("%-100s %8.4f ms %8.1f GB/s %.3f" % (r["kernel"][:100], r["ms"], r["gbs"], r["frac_of_measured_hbm_peak"]))
PY
tail -3 gpurun_out/r2g_bench_kernels.err
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:'spmv_stream' -c 8 -f -o gpurun_out/r2g_spmv python scripts/sp_only.py once > gpurun_out/r2g_ncu_spmv.out 2>&1; echo "ncu spmv rc=$?"; tail -2 gpurun_out/r2g_ncu_spmv.out
for route in fused stock; do
  timeout 600 python bench.py --workload c2_qp_n8192_m8192_p1024 --route $route --steps 200 --no-cpu-baseline --no-parity > gpurun_out/r2g_bench_c2_$route.json 2> gpurun_out/r2g_bench_c2_$route.err; echo "bench c2 $route rc=$?: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2g_bench_c2_$route.json) $(grep -o '"e2e": {"value": [0-9.]*' gpurun_out/r2g_bench_c2_$route.json)"; tail -2 gpurun_out/r2g_bench_c2_$route.err
  timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/r2g_launches_c2_$route.csv python bench.py --workload c2_qp_n8192_m8192_p1024 --route $route --steps 10 --warmup 10 --repeats 1 --no-cpu-baseline --no-parity > gpurun_out/r2g_ncu_c2_$route.out 2>&1; echo "ncu launches c2 $route rc=$?"; wc -l gpurun_out/r2g_launches_c2_$route.csv
done
timeout 600 python bench.py --workload c4_sdp_psd512_A131328x1024 --steps 100 --no-cpu-baseline --no-parity > gpurun_out/r2g_bench_c4.json 2> gpurun_out/r2g_bench_c4.err; echo "bench c4 rc=$?: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2g_bench_c4.json)"
timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 700 --csv --log-file gpurun_out/r2g_launches_c4.csv python bench.py --workload c4_sdp_psd512_A131328x1024 --steps 5 --warmup 5 --repeats 1 --no-cpu-baseline --no-parity > gpurun_out/r2g_ncu_c4.out 2>&1; echo "ncu launches c4 rc=$?"; wc -l gpurun_out/r2g_launches_c4.csv
timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 300 --csv --log-file gpurun_out/r2g_launches_c3.csv python bench.py --steps 5 --warmup 5 --repeats 1 --no-cpu-baseline --no-parity > gpurun_out/r2g_ncu_c3.out 2>&1; echo "ncu launches c3 rc=$?"; wc -l gpurun_out/r2g_launches_c3.csv
timeout 600 python bench.py > gpurun_out/r2g_bench_c3_default.json 2> gpurun_out/r2g_bench_c3_default.err; echo "bench c3 default rc=$?"; cut -c1-300 gpurun_out/r2g_bench_c3_default.json
du -sh gpurun_out
