#!/bin/bash
# Round 2, call I (gpurun --gpus N): sharded-operator parity tests (peer path and NCCL path), bench at N and 1 on the same box.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${NGPU:-2}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout -k 5 300 python -m pytest tests/test_gemv_gpu.py tests/test_solver_gpu.py -q -m gpu -k "transform_sp or scalar_prefetch" > gpurun_out/r2i_pytest_sp.log 2>&1; echo "pytest sp/prefetch rc=$?"; tail -4 gpurun_out/r2i_pytest_sp.log | cut -c1-300
timeout -k 10 1500 python -m pytest tests/test_dist_gpu.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/r2i_test_dist_n$N.log 2>&1; echo "== dist tests rc=$?"; tail -25 gpurun_out/r2i_test_dist_n$N.log | cut -c1-300
run_bench() {  # world, tag, env, extra args
  W=$1; TAG=$2; ENVV=$3; shift 3
  if [ "$W" = 1 ]; then
    env $ENVV timeout -k 10 400 python bench.py --gpus 1 "$@" > gpurun_out/r2i_bench_${TAG}.json 2> gpurun_out/r2i_bench_${TAG}.err
  else
    env $ENVV timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port $((29700+W)) \
       bench.py --gpus $W "$@" > gpurun_out/r2i_bench_${TAG}.json 2> gpurun_out/r2i_bench_${TAG}.err
  fi
  echo "== bench $TAG rc=$?: $(grep -o '"value": [0-9.]*' gpurun_out/r2i_bench_${TAG}.json | head -1) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2i_bench_${TAG}.json) $(grep -o '"e2e": {"value": [0-9.]*' gpurun_out/r2i_bench_${TAG}.json)"; tail -1 gpurun_out/r2i_bench_${TAG}.err | cut -c1-200
}
for W in ${WORLDS:-$N}; do
  run_bench $W c3_n${W}_driver X=1 --steps 20 --warmup 5
  run_bench $W c3_n${W} X=1 --steps 100 --warmup 10 --no-cpu-baseline --no-parity
  run_bench $W c3_n${W}_nccl TB_P2P=0 --steps 100 --warmup 10 --no-cpu-baseline --no-parity
  run_bench $W c3_n${W}_nospecx TB_P2P_SPEC=0 --steps 100 --warmup 10 --no-cpu-baseline --no-parity
done
run_bench 1 c3_n1 X=1 --steps 100 --warmup 10 --no-cpu-baseline --no-parity
du -sh gpurun_out
