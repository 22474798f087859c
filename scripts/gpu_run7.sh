#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_models.py -m gpu -x -q --timeout 300 --tb=short 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
rm -f gpurun_out/head_trace.txt*
PV2_TRACE=gpurun_out/head_trace.txt timeout 300 python bench_head.py --batches 16 --sizes 352 --iters 30 > gpurun_out/head_trace.log 2>&1
tail -1 gpurun_out/head_trace.log | cut -c1-120
head -12 gpurun_out/head_trace.txt
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench7.log 2>&1; tail -1 gpurun_out/bench7.log | cut -c1-300
