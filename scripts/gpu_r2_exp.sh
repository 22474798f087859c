#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py -x -q -m gpu -k "train_step or step_host or TrainStep or trainstep" 2>&1 | grep -E "^E   |^tests/|passed|failed|^FAILED" | head -20 | cut -c1-300
