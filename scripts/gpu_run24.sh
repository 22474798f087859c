#!/bin/bash
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 50 --tb=short -k "lowres" 2>&1 | tail -6 > gpurun_out/pytest_lowres24.log; tail -3 gpurun_out/pytest_lowres24.log
timeout 40 python __graft_entry__.py --smoke 2>&1 | tail -1
