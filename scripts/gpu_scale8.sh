#!/bin/bash
# 8-GPU scaling run (trimmed for GPU budget): C3 at N=4 and N=8, C5 at N=8.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run_bench() {  # world, tag, extra args
  W=$1; TAG=$2; shift 2
  timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port $((29700+W)) \
     bench.py --gpus $W --steps 100 --warmup 10 "$@" > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
  echo "== bench $TAG exit $?"; grep -h '^{' gpurun_out/bench_${TAG}.json | head -c 700; echo; tail -2 gpurun_out/bench_${TAG}.err
}
run_bench 8 c3_n8
run_bench 4 c3_n4
run_bench 8 c5_n8 --workload c5_lp_A262144x65536
