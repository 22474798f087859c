#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
  timeout 200 python bench_head.py --batches 16 --sizes 352 --iters 50 > gpurun_out/head_r$i.log 2>&1
  echo "run $i: $(tail -1 gpurun_out/head_r$i.log | cut -c1-70)"
done
