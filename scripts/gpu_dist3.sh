#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_dist_gpu.py -m gpu -q -x --no-header -p no:cacheprovider -k "sharded_operator" 2>&1 | tail -4
timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29661 \
   bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/bench_n2_final.json 2> gpurun_out/bench_n2_final.err
echo "== bench N=2 exit $?: $(grep -o '"value": [0-9.]*' gpurun_out/bench_n2_final.json | head -1) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_n2_final.json) $(grep -o '"gpu_launches": [0-9]*' gpurun_out/bench_n2_final.json)"
