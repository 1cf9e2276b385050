#!/bin/bash
# Round 2, call M (1 GPU): PSD sign-iteration schedules (error + time), C2 stock route with set_sqrt on the device, final spmv ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/r2m_psd_schedule.jsonl
for sch in 10,8 10,6 10,5 10,4 9,6 9,5 8,6; do
  TB_PSD_STEPS=$sch timeout 200 python scripts/psd_schedule.py >> gpurun_out/r2m_psd_schedule.jsonl 2> gpurun_out/r2m_psd_schedule.err || echo "schedule $sch failed"
done
python - <<'PY'
import json
for ln in open("gpurun_out/r2m_psd_schedule.jsonl"):
    d = json.loads(ln)
    print(d["schedule"], "ms %.3f" % d["ms_per_projection_k512"], " ".join("%s=%.1e" % (k[4:], v) for k, v in d.items() if k.startswith("err_")))
PY
tail -3 gpurun_out/r2m_psd_schedule.err
timeout 900 python bench.py --workload c2_qp_n8192_m8192_p1024 --route stock --steps 200 --no-cpu-baseline > gpurun_out/r2m_bench_c2_stock.json 2> gpurun_out/r2m_bench_c2_stock.err; echo "bench c2 stock rc=$?: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2m_bench_c2_stock.json) $(grep -o '"problem_construction_s": [0-9.]*' gpurun_out/r2m_bench_c2_stock.json) $(grep -o '"parity": {"pass": [a-z]*' gpurun_out/r2m_bench_c2_stock.json)"; tail -3 gpurun_out/r2m_bench_c2_stock.err
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:'spmv_stream' -c 8 -f -o gpurun_out/r2m_spmv python scripts/sp_only.py once > gpurun_out/r2m_ncu_spmv.out 2>&1; echo "ncu spmv rc=$?"; tail -2 gpurun_out/r2m_ncu_spmv.out
