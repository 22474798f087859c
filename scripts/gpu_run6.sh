#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/head_trace.txt*
PV2_TRACE=gpurun_out/head_trace.txt timeout 300 python bench_head.py --batches 16 --sizes 352 --iters 30 > gpurun_out/head_trace.log 2>&1
tail -1 gpurun_out/head_trace.log | cut -c1-120
head -30 gpurun_out/head_trace.txt
ls -la gpurun_out/head_trace.txt*
