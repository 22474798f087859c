#!/bin/bash
# Round check on one B200: GPU parity suite, smoke, bench line, head microbench, ncu launch list of the bench command.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 --tb=short 2>&1 | tail -25 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-1500
timeout 300 python bench_head.py --batches 16 --sizes 352 --iters 30 > gpurun_out/head.log 2>&1; tail -1 gpurun_out/head.log | cut -c1-400
timeout 600 python bench_head.py --batches 16 --sizes 352 --iters 10 --kernels --out gpurun_out/head_kernels.jsonl > gpurun_out/head_kernels.log 2>&1; tail -30 gpurun_out/head_kernels.log | cut -c1-220
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
python profiles/launch_summary.py gpurun_out/launches_bench.csv --last-fraction 0.3 2>&1 | head -40
