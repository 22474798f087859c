#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $1"; shift; env "$@" timeout 200 python bench_head.py --batches 16 --sizes 352 --iters 40 2>&1 | grep ms_graph | cut -c1-110; }
run late_trigger A=1
run cps2 PV2_CONV_CPS=2
run nopdl PV2_PDL=0
run v1 PV2_CONV_V1=1
timeout 900 python -m pytest tests/ -x -q -m gpu -k "bench_config" -s 2>&1 | grep -E "passed|failed|logits max|Error|mask agreement" | cut -c1-400
timeout 900 python -m pytest tests/ -x -q -m gpu -k "not bench_config" 2>&1 | tail -5
PV2_TRACE=gpurun_out/r2_timeline_v2.txt timeout 300 python bench_head.py --batches 16 --sizes 352 --iters 20 > gpurun_out/r2_trace.log 2>&1; echo "trace rc=$?"
