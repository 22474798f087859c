#!/bin/bash
timeout 900 python -m pytest tests/ -q -m gpu -k "decoder_class" 2>&1 | grep -E "^E  |passed|failed" | head -20 | cut -c1-500
for i in 1 2 3; do timeout 900 python -m pytest tests/ -x -q -m gpu -k "bench_config and fp32" 2>&1 | grep -E "^E  |passed|failed" | head -6 | cut -c1-400; done
