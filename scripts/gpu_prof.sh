#!/bin/bash
# ncu evidence for the bench command (B200_PROFILING.md): launch list with per-launch device time, and one
# --set full capture of the dominant kernel.  Numbers printed under ncu are never bench values.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline ${PROF_BENCH_ARGS}"
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.out 2>&1
echo "== ncu launches exit $?"; tail -3 gpurun_out/ncu_launches.out; wc -l gpurun_out/launches.csv
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 8 -c 3 -f -o gpurun_out/prof_stream $CMD > gpurun_out/ncu_full.out 2>&1
echo "== ncu full exit $?"; tail -3 gpurun_out/ncu_full.out; ls -la gpurun_out/*.ncu-rep
