#!/bin/bash
# Round-2 end validation on one B200: GPU tests as the driver runs them, smoke, the bench lines, ncu launch list + full capture
# of the streaming kernel, kernel micro-benchmarks.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2z_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2z_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2z_bench_c3_driver.json 2> gpurun_out/r2z_bench_c3_driver.err; echo "bench driver-flags rc=$?"; cut -c1-260 gpurun_out/r2z_bench_c3_driver.json; tail -2 gpurun_out/r2z_bench_c3_driver.err
timeout 900 python bench.py > gpurun_out/r2z_bench_c3_default.json 2> gpurun_out/r2z_bench_c3_default.err; echo "bench default rc=$?"; cut -c1-260 gpurun_out/r2z_bench_c3_default.json
for spec in "c2_qp_n8192_m8192_p1024 fused" "c4_sdp_psd512_A131328x1024 fused"; do
  set -- $spec
  timeout 600 python bench.py --workload $1 --route $2 --steps 200 --no-cpu-baseline --parity-k 100 > gpurun_out/r2z_bench_$1_$2.json 2> gpurun_out/r2z_bench_$1_$2.err; echo "bench $1 $2 rc=$?: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2z_bench_$1_$2.json) $(grep -o '"e2e": {"value": [0-9.]*' gpurun_out/r2z_bench_$1_$2.json) $(grep -o '"parity": {"pass": [a-z]*' gpurun_out/r2z_bench_$1_$2.json)"
done
timeout 600 python bench.py --dtype f64 --steps 100 --no-cpu-baseline --no-parity > gpurun_out/r2z_bench_c3_f64.json 2> gpurun_out/r2z_bench_c3_f64.err; echo "bench f64 rc=$?: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2z_bench_c3_f64.json)"
timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv --log-file gpurun_out/r2z_launches_c3.csv python bench.py --steps 5 --warmup 5 --repeats 1 --no-cpu-baseline --no-parity > gpurun_out/r2z_ncu_launches.out 2>&1
echo "== ncu launches exit $?"; wc -l gpurun_out/r2z_launches_c3.csv
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 6 -c 4 -f -o gpurun_out/r2z_stream python bench.py --steps 3 --warmup 3 --repeats 1 --no-cpu-baseline --no-parity > gpurun_out/r2z_ncu_full.out 2>&1
echo "== ncu full exit $?"; ls -la gpurun_out/r2z_stream.ncu-rep
timeout -k 5 400 python scripts/bench_kernels.py > gpurun_out/r2z_bench_kernels.json 2> gpurun_out/r2z_bench_kernels.err; echo "bench_kernels rc=$?"
du -sh gpurun_out
