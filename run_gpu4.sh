#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_conv.py -q --timeout 120 2>&1 | tail -40 > gpurun_out/pytest_conv.log
tail -40 gpurun_out/pytest_conv.log
