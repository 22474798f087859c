#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_infer_tail.py -m gpu -x -q --timeout 200 --tb=short 2>&1 | tail -8
timeout 300 python bench.py --mode infer --steps 30 --warmup 5 > gpurun_out/bench_infer.log 2>&1; tail -1 gpurun_out/bench_infer.log | cut -c1-1200
