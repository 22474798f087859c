#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/head_trace*.txt
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_models.py -m gpu -x -q --timeout 300 --tb=short 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench_head.py --batches 16 --sizes 352 --iters 30 > gpurun_out/head_A.log 2>&1; tail -1 gpurun_out/head_A.log | cut -c1-200
PV2_PDL=0 PV2_TRACE=gpurun_out/head_trace_nopdl.txt timeout 300 python bench_head.py --batches 16 --sizes 352 --iters 30 > gpurun_out/head_B.log 2>&1; tail -1 gpurun_out/head_B.log | cut -c1-200
head -36 gpurun_out/head_trace_nopdl.txt
