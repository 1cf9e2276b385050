#!/bin/bash
# Round 2, call U (1 GPU): a finalize with many partials counts as a long op (wide program) - A/B on C2 / C4 / C3
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for w in 262144 1000000000; do
  TB_VP_FINALIZE_WIDE=$w timeout 300 python bench.py --workload c2_qp_n8192_m8192_p1024 --steps 200 --no-cpu-baseline --no-parity ${TL:+--timeline gpurun_out/r2u_timeline_c2_$w.txt} > gpurun_out/r2u_bench_c2_$w.json 2> gpurun_out/r2u_bench_c2_$w.err; echo "c2 wide>$w rc=$?: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2u_bench_c2_$w.json) $(grep -o '"gpu_launches": [0-9]*' gpurun_out/r2u_bench_c2_$w.json)"
  TB_VP_FINALIZE_WIDE=$w timeout 300 python bench.py --workload c4_sdp_psd512_A131328x1024 --steps 100 --no-cpu-baseline --no-parity > gpurun_out/r2u_bench_c4_$w.json 2> gpurun_out/r2u_bench_c4_$w.err; echo "c4 wide>$w rc=$?: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2u_bench_c4_$w.json)"
  TB_VP_FINALIZE_WIDE=$w timeout 300 python bench.py --workload socp_small_128x64_A8192x4096 --steps 200 --no-cpu-baseline --no-parity > gpurun_out/r2u_bench_small_$w.json 2> gpurun_out/r2u_bench_small_$w.err; echo "socp_small wide>$w rc=$?: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2u_bench_small_$w.json)"
done
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-parity > gpurun_out/r2u_bench_c3.json 2> gpurun_out/r2u_bench_c3.err; echo "c3 rc=$?: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2u_bench_c3.json)"
