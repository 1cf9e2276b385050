#!/bin/bash
# Round 2 (gpurun --gpus 8): config C5 at full size - the parity test (68.7 GB of A row-sharded over 8 GPUs) and the bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_dist_gpu.py -m gpu -q --no-header -p no:cacheprovider -k "c5" > gpurun_out/r2c5_test.log 2>&1; echo "== c5 test rc=$?"; tail -5 gpurun_out/r2c5_test.log | cut -c1-300
timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29788 bench.py --gpus 8 --workload c5_lp_A262144x65536 --steps 50 --warmup 5 --no-cpu-baseline --no-parity > gpurun_out/r2c5_bench_n8.json 2> gpurun_out/r2c5_bench_n8.err; echo "== bench c5 rc=$?: $(grep -o '"value": [0-9.]*' gpurun_out/r2c5_bench_n8.json | head -1) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2c5_bench_n8.json)"; tail -2 gpurun_out/r2c5_bench_n8.err | cut -c1-200
