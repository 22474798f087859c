#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu_info.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --tb=short 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-2500
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step_r1b.csv python profiles/prof_kernels.py step > gpurun_out/ncu_step.log 2>&1; wc -l gpurun_out/launches_step_r1b.csv
