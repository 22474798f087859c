#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/ -x -q -m gpu --timeout 120 2>&1 | grep -E "^E   |^tests/|passed|failed|^FAILED|Timeout" | head -12 | cut -c1-300
for f in 1 0 1 0; do
  echo "fused $f"; PV2_BN_BWD_FUSED=$f timeout 120 python bench_head.py --batches 16 --sizes 352 --iters 60 2>&1 | grep -E "ms_graph" | cut -c1-100,185-215
done
PV2_TRACE=gpurun_out/r2_timeline_fused.txt timeout 120 python bench_head.py --batches 16 --sizes 352 --iters 20 > /dev/null 2>&1; head -14 gpurun_out/r2_timeline_fused.txt | cut -c1-100; rm -f gpurun_out/r2_timeline_fused.txt.chrome.json
