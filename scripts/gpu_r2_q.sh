#!/bin/bash
# Round 2, call Q (gpurun --gpus N): sharded parity tests after the launch-count work, bench at the listed worlds (and 1) on one box
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${NGPU:-2}
if [ "${SKIP_TESTS:-0}" != 1 ]; then
  timeout -k 10 900 python -m pytest tests/test_dist_gpu.py -m gpu -q --no-header -p no:cacheprovider -k "sharded" > gpurun_out/r2q_test_dist_n$N.log 2>&1; echo "== dist tests rc=$?"; tail -5 gpurun_out/r2q_test_dist_n$N.log | cut -c1-300
fi
run_bench() {  # world, tag, extra args
  W=$1; TAG=$2; shift 2
  if [ "$W" = 1 ]; then
    timeout -k 10 400 python bench.py --gpus 1 "$@" > gpurun_out/r2q_bench_${TAG}.json 2> gpurun_out/r2q_bench_${TAG}.err
  else
    timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port $((29700+W)) \
       bench.py --gpus $W "$@" > gpurun_out/r2q_bench_${TAG}.json 2> gpurun_out/r2q_bench_${TAG}.err
  fi
  echo "== bench $TAG rc=$?: $(grep -o '"value": [0-9.]*' gpurun_out/r2q_bench_${TAG}.json | head -1) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2q_bench_${TAG}.json) $(grep -o '"e2e": {"value": [0-9.]*' gpurun_out/r2q_bench_${TAG}.json) $(grep -o '"parity": {"pass": [a-z]*' gpurun_out/r2q_bench_${TAG}.json)"; tail -1 gpurun_out/r2q_bench_${TAG}.err | cut -c1-200
}
for W in ${WORLDS:-$N}; do
  run_bench $W c3_n${W}_driver --steps 20 --warmup 5 ${TL:+--timeline gpurun_out/r2q_timeline_c3_n$W.txt}
  run_bench $W c3_n${W} --steps 100 --warmup 10 --no-cpu-baseline --no-parity
done
run_bench 1 c3_n1_same_box --steps 100 --warmup 10 --no-cpu-baseline --no-parity
