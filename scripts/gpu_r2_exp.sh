#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench rc=$?"; tail -1 gpurun_out/r2_bench_default.json | cut -c1-300
timeout 300 python bench_head.py --batches 16 --sizes 352 --iters 40 --kernels --kernels-at 16x352 --out gpurun_out/r2_head_kernels_16x352.jsonl > gpurun_out/r2_hk16.log 2>&1; echo "kernels rc=$?"; grep ms_graph gpurun_out/r2_hk16.log | cut -c1-110
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1 | cut -c1-200
