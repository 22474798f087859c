#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/ -x -q -m gpu --timeout 200 2>&1 | grep -E "^E   |passed|failed|^FAILED|Timeout" | head -8 | cut -c1-250
for f in 1 2; do
  timeout 120 python bench_head.py --batches 16 --sizes 352 --iters 60 2>&1 | grep -E "ms_graph" | cut -c1-100
done
