#!/bin/bash
# ncu evidence for config C4: launch list of a short bench run + full capture of the batched tcgen05 GEMM.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --workload c4_sdp_psd512_A131328x1024 --steps 2 --warmup 3 --no-cpu-baseline"
timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 700 --csv --log-file gpurun_out/launches_c4.csv $CMD > gpurun_out/ncu_launches_c4.out 2>&1
echo "== ncu launches exit $?"; wc -l gpurun_out/launches_c4.csv
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:symm_gemm_tc -s 300 -c 2 -f -o gpurun_out/prof_psd_tc_batch $CMD > gpurun_out/ncu_full_c4.out 2>&1
echo "== ncu full exit $?"; ls -la gpurun_out/prof_psd_tc_batch.ncu-rep
