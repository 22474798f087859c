#!/bin/bash
# BASELINE config 5 on one 8-GPU box: the whole 7 x 5 (batch x size) head sweep, one replica per GPU (the head has no collective:
# per-GPU numbers + min / median over GPUs), the PVT head on a subset, and the N = 8 bench line.  gpurun --gpus 8.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench_head.py --batches 1,2,4,8,16,32,64 --sizes 256,352,448,512,704 --iters 20 --out gpurun_out/r2_head_sweep_8gpu.jsonl > gpurun_out/r2_head_sweep_8gpu.log 2>&1; echo "sweep rc=$?"
timeout 300 $TR bench_head.py --backbone pvt --batches 1,4,16,64 --sizes 256,352,704 --iters 20 --out gpurun_out/r2_head_sweep_pvt_8gpu.jsonl > gpurun_out/r2_head_sweep_pvt_8gpu.log 2>&1; echo "pvt sweep rc=$?"
timeout 400 $TR bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "bench n8 rc=$?"
tail -1 gpurun_out/r2_bench_n8.json | cut -c1-600
python - <<'PY'
import json
for f in ("gpurun_out/r2_head_sweep_8gpu.jsonl", "gpurun_out/r2_head_sweep_pvt_8gpu.jsonl"):
    try:
        for l in open(f):
            r = json.loads(l)
            if "ms_graph" in r:
                print(f"B={r['B']:3d} S={r['S']:4d} {r['ms_graph']:7.3f} ms (min {r.get('ms_graph_min_over_gpus', 0):7.3f}) {r['images_per_s']:9.1f} img/s/GPU")
    except Exception as e:
        print(f, e)
PY
