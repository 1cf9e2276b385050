#!/bin/bash
# 8-GPU check: sharded tests (world 4 peer path, C5 full size at world 8), then bench C3 at N=4 and N=8 and C5 at N=8.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
nvidia-smi topo -m 2>/dev/null | head -12 > gpurun_out/topo8.txt
timeout -k 10 1200 python -m pytest tests/test_dist_gpu.py -m gpu -q -x --no-header -p no:cacheprovider -k "c5 or 1" 2>&1 | tail -40 > gpurun_out/test_dist8.log
echo "== dist tests exit ${PIPESTATUS[0]}"; tail -15 gpurun_out/test_dist8.log
run_bench() {  # world, tag, extra args
  W=$1; TAG=$2; shift 2
  timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port $((29700+W)) \
     bench.py --gpus $W --steps 100 --warmup 10 "$@" > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
  echo "== bench $TAG exit $?"; grep -h '^{' gpurun_out/bench_${TAG}.json | head -c 1500; echo; tail -2 gpurun_out/bench_${TAG}.err
}
run_bench 4 c3_n4
run_bench 8 c3_n8
TB_P2P=0 run_bench 8 c3_n8_nccl
run_bench 8 c5_n8 --workload c5_lp_A262144x65536
