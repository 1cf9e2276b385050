#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_dist_gpu.py -m gpu -q -x --no-header -p no:cacheprovider -k "sharded_operator" 2>&1 | tail -8
for PAIR in 1 0; do
TB_P2P_PAIR=$PAIR timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29650+PAIR)) \
   bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/bench_n2_pair$PAIR.json 2> gpurun_out/bench_n2_pair$PAIR.err
echo "== bench N=2 pair=$PAIR exit $?: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_n2_pair$PAIR.json) $(grep -o '"gpu_launches": [0-9]*' gpurun_out/bench_n2_pair$PAIR.json)"; tail -2 gpurun_out/bench_n2_pair$PAIR.err | cut -c1-200
done
