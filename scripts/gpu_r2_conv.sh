#!/bin/bash
# round 2: persistent conv kernel -- parity, per-kernel table, head step + timeline
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -x -q -m gpu -s -k "bench_config or conv or models or ops" > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -8 gpurun_out/r2_pytest_gpu.log
timeout 300 python bench_head.py --batches 16 --sizes 352 --iters 30 --kernels --out gpurun_out/r2_head_kernels_v2.jsonl > gpurun_out/r2_head_kernels_v2.log 2>&1; echo "bench v2 rc=$?"
grep -E '"kernel": "(conv|struct)|ms_graph' gpurun_out/r2_head_kernels_v2.log | cut -c1-200
PV2_TRACE=gpurun_out/r2_timeline_v2.txt timeout 300 python bench_head.py --batches 16 --sizes 352 --iters 20 > gpurun_out/r2_trace.log 2>&1; echo "trace rc=$?"
head -16 gpurun_out/r2_timeline_v2.txt
