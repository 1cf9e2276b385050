#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run_bench() {  # world, tag, extra args
  W=$1; TAG=$2; shift 2
  timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port $((29700+W)) \
     bench.py --gpus $W --steps 100 --warmup 10 "$@" > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
  echo "== bench $TAG exit $?: $(grep -o '"value": [0-9.]*' gpurun_out/bench_${TAG}.json | head -1) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_${TAG}.json)"; tail -1 gpurun_out/bench_${TAG}.err | cut -c1-160
}
run_bench 8 c3_n8
run_bench 4 c3_n4
