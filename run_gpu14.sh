#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 --tb=short 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; tail -30 gpurun_out/pytest_gpu.log
timeout 300 python bench_head.py --batches 16 --sizes 352 --iters 30 > gpurun_out/head_A.log 2>&1; tail -1 gpurun_out/head_A.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_head.csv python profiles/prof_kernels.py head > gpurun_out/ncu_head.log 2>&1; wc -l gpurun_out/launches_head.csv
