#!/bin/bash
# refresh of the per-kernel tables + CUPTI timeline of the head step (profiles/r2_head_kernels_*.jsonl, r2_head_timeline.txt) and the GPU suite
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | grep -E "^E   *Assert|^E  |^tests/|passed|failed|^FAILED" | head -12 | cut -c1-300
rm -f gpurun_out/r2_timeline.txt*
PV2_TRACE=gpurun_out/r2_timeline.txt timeout 300 python bench_head.py --batches 16 --sizes 352 --iters 40 --kernels --kernels-at 16x352 --out gpurun_out/r2_head_kernels_16x352.jsonl > gpurun_out/r2_hk16.log 2>&1; echo "kernels 16x352 rc=$?"
grep ms_graph gpurun_out/r2_hk16.log | cut -c1-200
timeout 400 python bench_head.py --batches 64 --sizes 704 --iters 10 --kernels --kernels-at 64x704 --out gpurun_out/r2_head_kernels_64x704.jsonl > gpurun_out/r2_hk64.log 2>&1; echo "kernels 64x704 rc=$?"
grep ms_graph gpurun_out/r2_hk64.log | cut -c1-200
head -12 gpurun_out/r2_timeline.txt | cut -c1-110
rm -f gpurun_out/r2_timeline.txt.chrome.json
