#!/bin/bash
# One short call (16 GPU-minutes left in the round): parity of the loss-from-low-res kernels and of the re-parallelised bilinear
# backward, their timings, the training bench with the fused loss off / on, one lean ncu capture; then as much of the full suite as fits.
mkdir -p gpurun_out /tmp/ncu
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout 240 python -m pytest tests/test_gpu_ops.py -m gpu -x -q --timeout 200 --tb=short -k "lowres or bilinear or structure_loss or interpolate or dsra_fuse" 2>&1 | tail -15 > gpurun_out/pytest_new.log; tail -3 gpurun_out/pytest_new.log; el tests-new
PV2_LOSS_LOWRES=1 timeout 240 python -m pytest tests/test_gpu_models.py tests/test_optim.py -m gpu -q --timeout 200 --tb=short -k "lowres or train_step or step_host" 2>&1 | tail -15 > gpurun_out/pytest_ts.log; tail -3 gpurun_out/pytest_ts.log; el tests-trainstep
timeout 150 python bench_head.py --batches 16 --sizes 352 --iters 50 --lowres-loss both --kernels --out gpurun_out/head_kernels_r19.jsonl > gpurun_out/head_kernels_r19.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/head_kernels_r19.jsonl'):
    r = json.loads(l)
    if 'ms_graph' in r: print(f"head B={r['B']} S={r['S']} lowres={r['loss_from_lowres']} graph {r['ms_graph']:.3f} ms launches {r['pv2_launches']}")
    elif r.get('bound') == 'hbm': print(f"{r['us']:8.2f} us {r['achieved_gbs']:8.1f} GB/s {100*r['frac_of_hbm_peak']:5.1f}%  {r['kernel']}")
PY
el bench-head
PV2_LOSS_LOWRES=1 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_lowres1.log 2>&1; tail -1 gpurun_out/bench_lowres1.log | cut -c1-330; el bench-lowres-1
PV2_LOSS_LOWRES=0 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_lowres0.log 2>&1; tail -1 gpurun_out/bench_lowres0.log | cut -c1-330; el bench-lowres-0
SEC="--section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section WarpStateStats --section ComputeWorkloadAnalysis"
timeout 150 ncu --clock-control none $SEC -k regex:'fused_kernel|lowres|bilinear|structure_loss_bwd' --launch-skip 7 -c 7 -f -o /tmp/ncu/prof_lowres python profiles/prof_kernels.py lowres > gpurun_out/prof_lowres.log 2>&1
python profiles/summarize_ncu.py /tmp/ncu/prof_lowres.ncu-rep > gpurun_out/ncu_lowres_summary.txt 2>&1; wc -l gpurun_out/ncu_lowres_summary.txt; el ncu
timeout 400 python -m pytest tests -m gpu -q --timeout 200 --tb=short --durations=8 > gpurun_out/pytest_gpu_full.log 2>&1; tail -3 gpurun_out/pytest_gpu_full.log; el full-suite
