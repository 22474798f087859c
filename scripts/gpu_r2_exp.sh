#!/bin/bash
# scratch: the A/B of the day (rewritten per experiment)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/ -x -q -m gpu --timeout 200 2>&1 | tail -3
