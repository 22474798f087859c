#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/head_trace*.txt
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 --tb=short 2>&1 | tail -25 > gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench_head.py --batches 16 --sizes 352 --iters 30 --kernels --out gpurun_out/head_kernels.jsonl > gpurun_out/head_kernels.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/head_kernels.jsonl'):
    r=json.loads(l)
    if 'kernel' in r: print(f"{r['us']:8.2f} us  {r.get('achieved_gbs', r.get('achieved_tflops')):9.1f} {'GB/s' if 'achieved_gbs' in r else 'TF/s'}  frac {r.get('frac_of_hbm_peak', r.get('frac_of_bf16_peak')):.3f}  {r['kernel']}")
    else: print({k:r[k] for k in ('B','S','ms_graph','pv2_launches')})
PY
tail -3 gpurun_out/head_kernels.log | cut -c1-300
