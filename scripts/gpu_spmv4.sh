#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:spmv_stream_kernel -s 2 -c 1 -f -o gpurun_out/prof_spmv_stream python scripts/sp_only.py > gpurun_out/ncu_spmv_stream.out 2>&1
echo "ncu full rc=$?"
