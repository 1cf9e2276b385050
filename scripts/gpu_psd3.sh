#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cone_eig_gpu.py tests/test_solver_gpu.py -x -q -m gpu 2>&1 | tail -5
for pr in 1 0; do
cat > /tmp/pp.py <<PY
import sys, os
sys.path.insert(0, os.getcwd())
from totsu_b200 import capi
capi.init(0); capi.check(capi.lib().tb_set_psd_pairing($pr))
sys.argv = ["bench.py", "--workload", "c4_sdp_psd512_A131328x1024", "--steps", "200", "--no-cpu-baseline"]
import bench; bench.main()
PY
timeout 600 python /tmp/pp.py > gpurun_out/bench_c4_pair$pr.json 2> gpurun_out/bench_c4_pair$pr.err; echo "c4 pairing=$pr: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_c4_pair$pr.json) $(grep -o '"last_residuals": [^]]*]' gpurun_out/bench_c4_pair$pr.json)"; tail -1 gpurun_out/bench_c4_pair$pr.err | cut -c1-200
done
