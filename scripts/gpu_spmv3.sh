#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemv_gpu.py -x -q -m gpu -k "transform_sp or sympack" 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none --csv --log-file gpurun_out/launches_spmv.csv python scripts/sp_only.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches_spmv.csv',errors='replace')))
h=next(i for i,r in enumerate(rows) if r and r[0]=="ID")
for r in rows[h+1:]:
    if 'spmv_stream_kernel' in r[4] and 'time' in r[-3]: print(r[4].split('(')[0][:40], r[-3], r[-1], r[-2])
PY


timeout 600 python bench.py --workload c2_qp_n8192_m8192_p1024 --route stock --steps 200 --no-cpu-baseline > gpurun_out/bench_c2_stock.json 2> gpurun_out/bench_c2_stock.err; echo "c2 stock rc=$?: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_c2_stock.json)"
