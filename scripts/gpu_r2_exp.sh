#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/ -x -q -m gpu --timeout 120 2>&1 | grep -E "^E   |^tests/|passed|failed|^FAILED|Timeout" | head -12 | cut -c1-300
for f in 1 0 1 0; do
  echo "tiled $f"; PV2_UP2_TILED=$f timeout 120 python bench_head.py --batches 16 --sizes 352 --iters 60 2>&1 | grep -E "ms_graph" | cut -c1-100
done
PV2_TRACE=gpurun_out/r2_timeline_t.txt timeout 120 python bench_head.py --batches 16 --sizes 352 --iters 20 > /dev/null 2>&1; grep -E "up2|# B" gpurun_out/r2_timeline_t.txt | cut -c1-100; rm -f gpurun_out/r2_timeline_t.txt.chrome.json
