#!/bin/bash
# both bench arms at N = 2 under torchrun, as the driver launches them (gpurun --gpus 2)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521"
timeout 400 $TR bench.py --impl reference --gpus 2 --steps 2 --warmup 1 --ref-seconds 20 > gpurun_out/r2_bench_reference_n2.json 2> gpurun_out/r2_bench_reference_n2.err; echo "ref n2 rc=$?"; grep impl gpurun_out/r2_bench_reference_n2.json | cut -c1-400
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench n2 rc=$?"; grep metric gpurun_out/r2_bench_n2.json | cut -c1-500
