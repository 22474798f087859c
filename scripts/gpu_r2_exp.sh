#!/bin/bash
mkdir -p gpurun_out /tmp/ncu
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "bilinear" 2>&1 | grep -E "^E   *Assert|^E  |^tests/|passed|failed|^FAILED" | head -12 | cut -c1-300
timeout 300 python bench_head.py --batches 16 --sizes 352 --iters 40 --kernels --kernels-at 16x352 2>&1 | grep -E "ms_graph|bilinear" | cut -c1-170
timeout 300 ncu --clock-control none --set full -k regex:'bilinear_bwd' -c 2 -f -o /tmp/ncu/prof_bil python profiles/prof_kernels.py r2 > gpurun_out/prof_bil.log 2>&1
python profiles/summarize_ncu.py /tmp/ncu/prof_bil.ncu-rep > gpurun_out/r2_ncu_bilbwd.txt 2>&1
ncu -i /tmp/ncu/prof_bil.ncu-rep --page source --csv --print-source cuda,sass 2>/dev/null | head -400 > gpurun_out/r2_ncu_bilbwd_source.csv
grep -E "duration|dram read|dram %|L2 bytes|issue slots|stall|occupancy|warp instr" gpurun_out/r2_ncu_bilbwd.txt | head -40
