#!/bin/bash
cd "$(dirname "$0")/.."
for w in socp_small_128x64_A8192x4096 c2_qp_n8192_m8192_p1024 c3_socp_1024x64_A65536x16384; do
  for mode in 1 2 3; do
    out=$(TB_VPROG_MODE=$mode timeout 300 python bench.py --workload $w --steps 200 --no-cpu-baseline 2>/dev/null)
    echo "$w mode=$mode: $(echo "$out" | grep -o '"ms_per_step": [0-9.]*')"
  done
  out=$(timeout 300 python bench.py --workload $w --steps 200 --no-cpu-baseline --vprog 0 2>/dev/null)
  echo "$w vprog off: $(echo "$out" | grep -o '"ms_per_step": [0-9.]*')"
done
