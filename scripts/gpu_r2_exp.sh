#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_models.py -x -q -m gpu --timeout 120 -k "ra_v1 or v1" 2>&1 | grep -E "^E   |^tests/|passed|failed|^FAILED|Timeout" | head -12 | cut -c1-300
for m in 2 1; do echo "flat mode $m"; PV2_RA_FLAT=$m timeout 300 python bench_head.py --batches 16 --sizes 352 --iters 40 --kernels --kernels-at 16x352 2>&1 | grep -E "ra_v1" | cut -c1-50,60-150; done
echo "64x704"; timeout 300 python bench_head.py --batches 64 --sizes 704 --iters 5 --kernels --kernels-at 64x704 2>&1 | grep -E "ra_v1" | cut -c1-50,60-150
