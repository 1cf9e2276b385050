#!/bin/bash
# ncu evidence for the ConePSD tensor-core path: launch list of one projection benchmark + full captures of the GEMM.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PSD_BENCH_JACOBI=0
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_psd.csv python scripts/bench_psd.py 512 2 > gpurun_out/ncu_launches_psd.out 2>&1
echo "== ncu launches exit $?"; tail -2 gpurun_out/ncu_launches_psd.out; wc -l gpurun_out/launches_psd.csv
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:symm_gemm_tc -s 60 -c 2 -f -o gpurun_out/prof_psd_tc python scripts/bench_psd.py 512 2 > gpurun_out/ncu_full_psd.out 2>&1
echo "== ncu full (split-K) exit $?"; tail -2 gpurun_out/ncu_full_psd.out
TB_TC_MAX_SPLITK=1 timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:symm_gemm_tc -s 60 -c 1 -f -o gpurun_out/prof_psd_tc_nosplit python scripts/bench_psd.py 512 2 > gpurun_out/ncu_full_psd1.out 2>&1
echo "== ncu full (no split) exit $?"; tail -2 gpurun_out/ncu_full_psd1.out
ls -la gpurun_out/*.ncu-rep
