#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | grep -E "^E   *Assert|^E  |^tests/|passed|failed|^FAILED" | head -12 | cut -c1-300
timeout 300 python bench_head.py --batches 16 --sizes 352 --iters 40 --kernels --kernels-at 16x352 2>&1 | grep -E "ms_graph|bilinear" | cut -c1-170
PV2_BIL_BWD2=0 timeout 300 python bench_head.py --batches 16 --sizes 352 --iters 40 2>&1 | grep -E "ms_graph" | cut -c1-160
