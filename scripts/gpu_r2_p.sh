#!/bin/bash
# Round 2, call P (1 GPU): prefetch riding in the trigger's program vs the separate kernel, C2 and C3, with timelines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for ride in 1 0; do
  TB_PF_IN_PROGRAM=$ride timeout 600 python bench.py --workload c2_qp_n8192_m8192_p1024 --steps 200 --no-cpu-baseline --no-parity --timeline gpurun_out/r2p_timeline_c2_ride$ride.txt > gpurun_out/r2p_bench_c2_ride$ride.json 2> gpurun_out/r2p_bench_c2_ride$ride.err; echo "bench c2 ride=$ride rc=$?: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2p_bench_c2_ride$ride.json) $(grep -o '"gpu_launches": [0-9]*' gpurun_out/r2p_bench_c2_ride$ride.json)"
  TB_PF_IN_PROGRAM=$ride timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-parity > gpurun_out/r2p_bench_c3_ride$ride.json 2> gpurun_out/r2p_bench_c3_ride$ride.err; echo "bench c3 ride=$ride rc=$?: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2p_bench_c3_ride$ride.json) $(grep -o '"gpu_launches": [0-9]*' gpurun_out/r2p_bench_c3_ride$ride.json)"
  TB_PF_IN_PROGRAM=$ride timeout 600 python bench.py --workload c4_sdp_psd512_A131328x1024 --steps 100 --no-cpu-baseline --no-parity > gpurun_out/r2p_bench_c4_ride$ride.json 2> gpurun_out/r2p_bench_c4_ride$ride.err; echo "bench c4 ride=$ride rc=$?: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2p_bench_c4_ride$ride.json)"
done
