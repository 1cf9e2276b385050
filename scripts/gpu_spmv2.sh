#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemv_gpu.py -x -q -m gpu -k "transform_sp or sympack" 2>&1 | tail -3
cat > /tmp/sp_only.py <<'PY'
import sys, os, math, numpy as np
sys.path.insert(0, os.getcwd())
from totsu_b200 import capi
capi.init(0); L = capi.lib(); dt = np.float32
for n in (8192, 16384):
    sp = capi.Buf(dtype=dt, length=n*(n+1)//2)
    capi.check(capi.fn("tb_fill_uniform", dt)(sp.view(), n*(n+1)//2, 1, 0, 1, dt(0.01)))
    x, y = capi.Buf(dtype=dt, length=n), capi.Buf(dtype=dt, length=n)
    x.upload(np.ones(n, dtype=dt))
    for _ in range(6):
        capi.check(capi.fn("tb_transform_sp", dt)(n, 1.0, sp.view(), x.view(), 0.0, y.view()))
    capi.check(L.tb_device_sync())
PY
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none --csv --log-file gpurun_out/launches_spmv.csv python /tmp/sp_only.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches_spmv.csv',errors='replace')))
h=next(i for i,r in enumerate(rows) if r and r[0]=="ID")
for r in rows[h+1:]:
    if 'spmv' in r[4]: print(r[4].split('(')[0][:40], r[-3], r[-1], r[-2])
PY
timeout 600 python scripts/bench_kernels.py 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin)
for r in d['rows']:
    if 'transform_sp' in r['kernel']: print(r['kernel'], r['ms'], r['gbs'], r['frac_of_measured_hbm_peak'])
"
