#!/bin/bash
# Last call of the round: full GPU suite on the final code first, then the per-kernel table, the knob A/B and the bench line.
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout 200 python -m pytest tests -m gpu -q --timeout 150 --tb=short > gpurun_out/pytest_gpu_full23.log 2>&1; tail -3 gpurun_out/pytest_gpu_full23.log; el full-suite
timeout 100 python bench_head.py --batches 16 --sizes 352 --iters 50 --lowres-loss both --kernels --out gpurun_out/head_kernels_r23.jsonl > gpurun_out/head_kernels_r23.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/head_kernels_r23.jsonl'):
    r = json.loads(l)
    if 'ms_graph' in r: print(f"head B={r['B']} S={r['S']} lowres={r['loss_from_lowres']} graph {r['ms_graph']:.3f} ms launches {r['pv2_launches']}")
    elif r.get('bound') == 'hbm': print(f"{r['us']:8.2f} us {r['achieved_gbs']:8.1f} GB/s {100*r['frac_of_hbm_peak']:5.1f}%  {r['kernel']}")
PY
el bench-head
timeout 100 python scripts/variants_bench.py > gpurun_out/variants23.log 2>&1; grep "bilinear" gpurun_out/variants23.log | tail -12; cp gpurun_out/variants.jsonl gpurun_out/variants23.jsonl; el variants
timeout 200 python bench.py > gpurun_out/bench23.log 2>&1; tail -1 gpurun_out/bench23.log | cut -c1-300; el bench
