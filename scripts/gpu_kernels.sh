#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python scripts/bench_kernels.py > gpurun_out/bench_kernels.json 2> gpurun_out/bench_kernels.err; echo "bench_kernels rc=$?"; tail -3 gpurun_out/bench_kernels.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_kernels.json'))
for r in d["rows"]: print("%-70s %8.4f ms %8.1f GB/s  %.2f" % (r["kernel"], r["ms"], r["gbs"], r["frac_of_measured_hbm_peak"]))
PY
# ncu: launch list of a C2 stock-route run (spmv, generic gemv, stock cones) and full captures of spmv / cone / vprog kernels
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_c2_stock.csv python bench.py --workload c2_qp_n8192_m8192_p1024 --route stock --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c2.out 2>&1
echo "== ncu launches c2 exit $?"; wc -l gpurun_out/launches_c2_stock.csv
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:spmv_kernel -s 6 -c 2 -f -o gpurun_out/prof_spmv python bench.py --workload c2_qp_n8192_m8192_p1024 --route stock --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_spmv.out 2>&1
echo "== ncu spmv exit $?"
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k 'regex:cone_kernel|vprog_kernel' -s 20 -c 4 -f -o gpurun_out/prof_cone_vprog python bench.py --workload socp_small_128x64_A8192x4096 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_cone.out 2>&1
echo "== ncu cone/vprog exit $?"
ls -la gpurun_out/*.ncu-rep
