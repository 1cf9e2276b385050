#!/bin/bash
# Multi-GPU check (gpurun --gpus N): sharded-operator tests on both collective paths, then bench at N with each.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${NGPU:-2}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
nvidia-smi topo -m 2>/dev/null | head -12 > gpurun_out/topo.txt
timeout -k 10 900 python -m pytest tests/test_dist_gpu.py -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/test_dist.log
echo "== dist tests exit ${PIPESTATUS[0]}"; tail -25 gpurun_out/test_dist.log
for P in 1 0; do
  for W in ${WORLDS:-$N}; do
    TB_P2P=$P timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port $((29600+W+10*P)) \
       bench.py --gpus $W --steps ${STEPS:-100} --warmup 10 > gpurun_out/bench_n${W}_p2p${P}.json 2> gpurun_out/bench_n${W}_p2p${P}.err
    echo "== bench N=$W p2p=$P exit $?"; head -c 1800 gpurun_out/bench_n${W}_p2p${P}.json; echo; tail -3 gpurun_out/bench_n${W}_p2p${P}.err
  done
done
