#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_models.py tests/test_gpu_conv.py -m gpu -x -q --timeout 300 --tb=short 2>&1 | tail -6
for i in 1 2; do
  timeout 200 python bench_head.py --batches 16 --sizes 352 --iters 50 > gpurun_out/head_r$i.log 2>&1
  echo "run $i: $(tail -1 gpurun_out/head_r$i.log | cut -c1-70)"
done
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench18.log 2>&1; tail -1 gpurun_out/bench18.log | cut -c1-900
