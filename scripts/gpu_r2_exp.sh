#!/bin/bash
mkdir -p gpurun_out
for p in 0 -1 0 -1; do
  echo "prio $p"; PV2_CHAIN_PRIORITY=$p timeout 300 python bench_head.py --batches 16 --sizes 352 --iters 60 2>&1 | grep -E "ms_graph" | cut -c1-100
done
