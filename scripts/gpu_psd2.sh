#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cone_eig_gpu.py tests/test_solver_gpu.py -x -q -m gpu -k "psd or sdp or eig or cone" 2>&1 | tail -5
PSD_BENCH_JACOBI=0 timeout 300 python scripts/bench_psd.py 512 20 > gpurun_out/bench_psd512.json 2>/dev/null; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_psd512.json'))
for k,v in d.items():
    if isinstance(v,dict): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items()})
PY
timeout 600 python bench.py --workload c4_sdp_psd512_A131328x1024 --steps 200 --no-cpu-baseline > gpurun_out/bench_c4_new.json 2>/dev/null; echo "c4: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_c4_new.json) $(grep -o '"last_residuals": [^]]*]' gpurun_out/bench_c4_new.json)"
