#!/bin/bash
# Round-end verification on one B200: GPU parity suite, smoke, bench lines (train / infer / reference arm), head microbench + sweep.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 --tb=short 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_final.log 2>&1; tail -1 gpurun_out/bench_final.log | cut -c1-2500
timeout 300 python bench.py --mode infer --steps 50 --warmup 5 > gpurun_out/bench_infer.log 2>&1; tail -1 gpurun_out/bench_infer.log | cut -c1-700
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-400
rm -f gpurun_out/head_trace.txt*
PV2_TRACE=gpurun_out/head_trace.txt timeout 300 python bench_head.py --batches 16 --sizes 352 --iters 50 --kernels --out gpurun_out/head_kernels_final.jsonl > gpurun_out/head_kernels_final.log 2>&1
grep ms_graph gpurun_out/head_kernels_final.log | cut -c1-120
timeout 600 python bench_head.py --batches 1,4,16,64 --sizes 256,352,704 --iters 20 --out gpurun_out/head_sweep.jsonl > gpurun_out/head_sweep.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/head_sweep.jsonl'):
    r = json.loads(l)
    if 'ms_graph' in r: print(f"B={r['B']:3d} S={r['S']:4d} graph {r['ms_graph']:7.3f} ms  eager {r['ms_eager']:8.2f} ms  {r['images_per_s']:9.1f} img/s  conv {r['conv_tflops_fwd_bwd']:7.1f} TFLOP/s")
PY
rm -f gpurun_out/head_trace.txt.chrome.json
