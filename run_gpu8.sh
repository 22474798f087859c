#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --tb=short 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-1800
timeout 600 ncu --set full --clock-control none --import-source on -k regex:structure_loss -s 3 -c 3 -f -o gpurun_out/prof_loss python profiles/prof_kernels.py loss > gpurun_out/ncu_loss.log 2>&1; tail -2 gpurun_out/ncu_loss.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_ -s 102 -c 102 -f -o gpurun_out/prof_conv python profiles/prof_kernels.py head > gpurun_out/ncu_conv.log 2>&1; tail -2 gpurun_out/ncu_conv.log
ls -la gpurun_out/*.ncu-rep
