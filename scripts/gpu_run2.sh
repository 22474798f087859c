#!/bin/bash
# New kernels' parity + conv co-residency A/B on one B200.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_optim.py tests/test_infer_tail.py tests/test_gpu_conv.py tests/test_gpu_models.py -m gpu -x -q --timeout 300 --tb=short 2>&1 | tail -30 > gpurun_out/pytest_new.log; tail -12 gpurun_out/pytest_new.log
for cfg in "dflt:" "c1:PV2_CONV_CTAS=1" "c2:PV2_CONV_CTAS=2" "c3:PV2_CONV_CTAS=3"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 200 python bench_head.py --batches 16 --sizes 352 --iters 30 > gpurun_out/head_$name.log 2>&1
  echo "$name [$envs]: $(tail -1 gpurun_out/head_$name.log | cut -c1-120)"
done
timeout 300 python bench_head.py --batches 16 --sizes 352 --iters 10 --kernels --out gpurun_out/head_kernels2.jsonl > gpurun_out/head_kernels2.log 2>&1
grep '"bound": "tensor"' gpurun_out/head_kernels2.jsonl | cut -c1-230
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench2.log 2>&1; tail -1 gpurun_out/bench2.log | cut -c1-1800
