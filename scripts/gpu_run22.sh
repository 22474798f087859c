#!/bin/bash
# Bilinear backward with tap tables / folded row groups, forward band heuristic, loss-forward prefetch modes: parity, A/B, per-kernel table,
# bench line, lean ncu capture of the touched kernels, full suite.
mkdir -p gpurun_out /tmp/ncu
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout 200 python -m pytest tests/test_gpu_ops.py -m gpu -x -q --timeout 200 --tb=short 2>&1 | tail -8 > gpurun_out/pytest_ops22.log; tail -2 gpurun_out/pytest_ops22.log; el ops-tests
timeout 200 python scripts/variants_bench.py > gpurun_out/variants22.log 2>&1; grep -v "^$" gpurun_out/variants22.log | tail -40; cp gpurun_out/variants.jsonl gpurun_out/variants22.jsonl; el variants
timeout 150 python bench_head.py --batches 16 --sizes 352 --iters 50 --lowres-loss both --kernels --out gpurun_out/head_kernels_r22.jsonl > gpurun_out/head_kernels_r22.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/head_kernels_r22.jsonl'):
    r = json.loads(l)
    if 'ms_graph' in r: print(f"head B={r['B']} S={r['S']} lowres={r['loss_from_lowres']} graph {r['ms_graph']:.3f} ms launches {r['pv2_launches']}")
    elif r.get('bound') == 'hbm': print(f"{r['us']:8.2f} us {r['achieved_gbs']:8.1f} GB/s {100*r['frac_of_hbm_peak']:5.1f}%  {r['kernel']}")
PY
el bench-head
timeout 300 python bench.py > gpurun_out/bench22.log 2>&1; tail -1 gpurun_out/bench22.log | cut -c1-420; el bench
SEC="--section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section WarpStateStats --section ComputeWorkloadAnalysis --section SchedulerStats --section InstructionStats"
timeout 150 ncu --clock-control none $SEC -k regex:'fused_kernel|bilinear' --launch-skip 4 -c 4 -f -o /tmp/ncu/prof_bil python profiles/prof_kernels.py lowres > gpurun_out/prof_bil.log 2>&1
python profiles/summarize_ncu.py /tmp/ncu/prof_bil.ncu-rep > gpurun_out/ncu_bil_summary.txt 2>&1; wc -l gpurun_out/ncu_bil_summary.txt; el ncu
timeout 300 python -m pytest tests -m gpu -q --timeout 200 --tb=short > gpurun_out/pytest_gpu_full22.log 2>&1; tail -3 gpurun_out/pytest_gpu_full22.log; el full-suite
