#!/bin/bash
mkdir -p gpurun_out
T="tests/test_gpu_conv.py tests/test_gpu_models.py"
timeout 900 python -m pytest $T -m gpu -x -q --timeout 300 --tb=short 2>&1 | tail -25 > gpurun_out/pytest_a.log; tail -3 gpurun_out/pytest_a.log
if ! grep -q " passed" gpurun_out/pytest_a.log || grep -q "failed\|error" gpurun_out/pytest_a.log; then
  PV2_CONV_PATCH=1 timeout 900 python -m pytest $T -m gpu -x -q --timeout 300 --tb=short 2>&1 | tail -25 > gpurun_out/pytest_b.log; echo "PATCH:"; tail -3 gpurun_out/pytest_b.log
  PV2_PDL=0 timeout 900 python -m pytest $T -m gpu -x -q --timeout 300 --tb=short 2>&1 | tail -25 > gpurun_out/pytest_c.log; echo "NOPDL:"; tail -3 gpurun_out/pytest_c.log
fi
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --tb=short 2>&1 | tail -25 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
for cfg in "A:" "B:PV2_PDL=0" "C:PV2_CONV_PATCH=1" "D:PV2_PDL=0 PV2_CONV_PATCH=1"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 300 python bench_head.py --batches 16 --sizes 352 --iters 30 > gpurun_out/head_$name.log 2>&1
  echo "$name [$envs]: $(tail -1 gpurun_out/head_$name.log | cut -c1-260)"
done
timeout 600 python bench_head.py --batches 16 --sizes 352 --iters 30 --kernels --out gpurun_out/head_kernels.jsonl > gpurun_out/head_kernels.log 2>&1; tail -22 gpurun_out/head_kernels.log | cut -c1-250
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-900
