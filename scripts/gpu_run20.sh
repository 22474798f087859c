#!/bin/bash
# Tuning knobs of the memory-bound kernels measured in one process (scripts/variants_bench.py), then the full GPU suite.
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout 200 python -m pytest tests/test_gpu_ops.py -m gpu -x -q --timeout 200 --tb=short 2>&1 | tail -8 > gpurun_out/pytest_ops20.log; tail -2 gpurun_out/pytest_ops20.log; el ops-tests
timeout 240 python scripts/variants_bench.py > gpurun_out/variants.log 2>&1; grep -v "^$" gpurun_out/variants.log | tail -40; el variants
timeout 300 python -m pytest tests -m gpu -q --timeout 200 --tb=short > gpurun_out/pytest_gpu_full20.log 2>&1; tail -3 gpurun_out/pytest_gpu_full20.log; el full-suite
