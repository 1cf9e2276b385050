#!/bin/bash
# Round-end validation on one B200: GPU tests, smoke, the bench lines, ncu launch list + full capture of the streaming kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default rc=$?"; cut -c1-300 gpurun_out/bench_default.json; tail -2 gpurun_out/bench_default.err
timeout 600 python bench.py --impl reference --steps 6 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench reference rc=$?"; cut -c1-200 gpurun_out/bench_reference.json
timeout 600 python bench.py --dtype f64 --steps 100 --no-cpu-baseline > gpurun_out/bench_c3_f64.json 2> gpurun_out/bench_c3_f64.err; echo "bench f64 rc=$?: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_c3_f64.json)"
for spec in "c2_qp_n8192_m8192_p1024 fused" "c2_qp_n8192_m8192_p1024 stock" "c4_sdp_psd512_A131328x1024 fused"; do
  set -- $spec
  timeout 600 python bench.py --workload $1 --route $2 --steps 200 > gpurun_out/bench_$1_$2.json 2> gpurun_out/bench_$1_$2.err; echo "bench $1 $2 rc=$?: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_$1_$2.json)"
done
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.out 2>&1
echo "== ncu launches exit $?"; wc -l gpurun_out/launches_bench.csv
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 6 -c 4 -f -o gpurun_out/prof_stream22 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.out 2>&1
echo "== ncu full exit $?"; ls -la gpurun_out/prof_stream22.ncu-rep
