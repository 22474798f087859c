#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2.log 2>&1; tail -1 gpurun_out/bench_n2.log | cut -c1-700
grep -i "error\|Traceback" gpurun_out/bench_n2.log | head -5
