#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_ops.py tests/test_gpu_models.py -m gpu -x -q --timeout 300 --tb=short 2>&1 | tail -6
timeout 300 python bench_head.py --batches 16 --sizes 352 --iters 30 --kernels --out gpurun_out/head_kernels4.jsonl > gpurun_out/head_kernels4.log 2>&1
grep '"bound": "hbm"' gpurun_out/head_kernels4.jsonl | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(f\"{r['us']:8.2f} us {r['achieved_gbs']:8.1f} GB/s {r['frac_of_hbm_peak']:.3f}  {r['kernel']}\")"
grep ms_graph gpurun_out/head_kernels4.log | cut -c1-80
