#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_models.py -m gpu -x -q --timeout 300 --tb=short 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
for cfg in "dflt:" "nowg:PV2_WGRAD_STREAMS=0"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 200 python bench_head.py --batches 16 --sizes 352 --iters 30 > gpurun_out/head_$name.log 2>&1
  echo "$name [$envs]: $(tail -1 gpurun_out/head_$name.log | cut -c1-120)"
done
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench4.log 2>&1; tail -1 gpurun_out/bench4.log | cut -c1-400
