#!/bin/bash
mkdir -p gpurun_out /tmp/ncu
run() { echo "== $1"; shift; env "$@" timeout 200 python bench_head.py --batches 16 --sizes 352 --iters 40 2>&1 | grep ms_graph | cut -c1-110; }
run base A=1
run maxgrid96 PV2_CONV_MAXGRID=96
run v1 PV2_CONV_V1=1
timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4
PV2_TRACE=gpurun_out/r2_timeline_v2.txt timeout 300 python bench_head.py --batches 16 --sizes 352 --iters 20 --kernels --out gpurun_out/r2_head_kernels_v2.jsonl > gpurun_out/r2_trace.log 2>&1; echo "trace rc=$?"
grep -E '"kernel": "(conv_fwd\+|struct)' gpurun_out/r2_trace.log | cut -c1-170
head -14 gpurun_out/r2_timeline_v2.txt
