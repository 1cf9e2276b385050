#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemv_gpu.py -x -q -m gpu -k "transform_sp or sympack" 2>&1 | tail -6
timeout 600 python scripts/bench_kernels.py > gpurun_out/bench_kernels.json 2> gpurun_out/bench_kernels.err; echo "bench_kernels rc=$?"; tail -3 gpurun_out/bench_kernels.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_kernels.json'))
for r in d["rows"]: print("%-70s %8.4f ms %8.1f GB/s  %.2f" % (r["kernel"], r["ms"], r["gbs"], r["frac_of_measured_hbm_peak"]))
PY
timeout 600 python bench.py --workload c2_qp_n8192_m8192_p1024 --route stock --steps 200 --no-cpu-baseline > gpurun_out/bench_c2_stock.json 2> gpurun_out/bench_c2_stock.err; echo "c2 stock rc=$?: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_c2_stock.json)"
