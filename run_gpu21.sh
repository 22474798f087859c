#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 --tb=short 2>&1 | tail -25 > gpurun_out/pytest_gpu.log; tail -8 gpurun_out/pytest_gpu.log
for cfg in "A:" "B:PV2_STREAMS=0"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 300 python bench_head.py --batches 16 --sizes 352 --iters 30 > gpurun_out/head_$name.log 2>&1
  echo "$name [$envs]: $(tail -1 gpurun_out/head_$name.log | cut -c1-160)"
done
timeout 600 python bench_head.py --batches 16 --sizes 352 --iters 10 --kernels --out gpurun_out/head_kernels.jsonl > gpurun_out/head_kernels.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/head_kernels.jsonl'):
    r=json.loads(l)
    if 'kernel' in r and r['bound']=='hbm': print(f"{r['us']:8.2f} us  {r.get('achieved_gbs'):9.1f} GB/s  frac {r.get('frac_of_hbm_peak'):.3f}  {r['kernel']}")
PY
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-400
