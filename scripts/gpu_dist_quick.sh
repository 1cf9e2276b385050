#!/bin/bash
# quick 2-GPU sanity: sharded operator / solver tests on the peer path, one short bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_dist_gpu.py -m gpu -q -x --no-header -p no:cacheprovider -k "sharded_operator and 1" 2>&1 | tail -3
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29671 \
   bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_n2_last.json 2> gpurun_out/bench_n2_last.err
echo "== bench N=2 exit $?: $(grep -o '"value": [0-9.]*' gpurun_out/bench_n2_last.json | head -1)"
