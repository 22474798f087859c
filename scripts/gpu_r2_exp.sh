#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | grep -E "^E   *Assert|^tests/|passed|failed|^FAILED" | head -8 | cut -c1-300
timeout 300 python bench_head.py --batches 16 --sizes 352 --iters 40 2>&1 | grep ms_graph | cut -c1-100
timeout 300 python bench_head.py --kernels-at 16x352 > gpurun_out/r2_head_kernels_16x352_b.jsonl 2>/dev/null
grep -E "up2|bilinear|conv" gpurun_out/r2_head_kernels_16x352_b.jsonl | cut -c1-220 | head -40
