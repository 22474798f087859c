#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_optim.py tests/test_gpu_bench_config.py -x -q -m gpu --timeout 200 2>&1 | grep -E "^E   |passed|failed|^FAILED|Timeout" | head -6 | cut -c1-250
for c in -1 100 -1 100; do
  echo "stream carveout $c"; PV2_CARVEOUT_STREAM=$c timeout 120 python bench_head.py --batches 16 --sizes 352 --iters 60 2>&1 | grep -E "ms_graph" | cut -c1-100
done
for c in -1 100; do
  echo "stream carveout $c (kernel rows)"; PV2_CARVEOUT_STREAM=$c timeout 200 python bench_head.py --batches 16 --sizes 352 --iters 10 --kernels --kernels-at 16x352 2>&1 | grep -E "prepared|loss bwd x4\"|8 final maps" | cut -c12-60,100-140
done
