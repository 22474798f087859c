#!/bin/bash
mkdir -p gpurun_out /tmp/ncu
timeout 45 ncu --clock-control none --set full -k regex:'bilinear_fwd' --launch-skip 1 -c 1 -f -o /tmp/ncu/prof_bilfwd python profiles/prof_kernels.py lowres > gpurun_out/prof_bilfwd.log 2>&1
python profiles/summarize_ncu.py /tmp/ncu/prof_bilfwd.ncu-rep > gpurun_out/ncu_bilfwd_summary.txt 2>&1; grep -c . gpurun_out/ncu_bilfwd_summary.txt
timeout 45 python bench.py --mode infer --steps 30 --warmup 5 > gpurun_out/bench_infer25.log 2>&1; tail -1 gpurun_out/bench_infer25.log | cut -c1-260
