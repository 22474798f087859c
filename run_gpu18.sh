#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/head_trace*.txt
PV2_NO_FUSED_STATS=1 PV2_PDL=0 PV2_TRACE=gpurun_out/head_trace_nofuse.txt timeout 300 python bench_head.py --batches 16 --sizes 352 --iters 30 > gpurun_out/head_B.log 2>&1; tail -1 gpurun_out/head_B.log | cut -c1-200
head -12 gpurun_out/head_trace_nofuse.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_fwd_kernel -s 120 -c 12 -f -o gpurun_out/prof_conv_r1c python profiles/prof_kernels.py head > gpurun_out/ncu_conv.log 2>&1; tail -2 gpurun_out/ncu_conv.log
