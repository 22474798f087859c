#!/bin/bash
mkdir -p gpurun_out
for b in 48 24 32 40; do
  echo "budget $b"; PV2_PAR_CTA_BUDGET=$b timeout 120 python bench_head.py --batches 16 --sizes 352 --iters 60 2>&1 | grep -E "ms_graph" | cut -c1-100
done
for nb in 148 96 74; do
  echo "budget 48 narrow $nb"; PV2_PAR_CTA_BUDGET=48 PV2_PAR_CTA_BUDGET_NARROW=$nb timeout 120 python bench_head.py --batches 16 --sizes 352 --iters 60 2>&1 | grep -E "ms_graph" | cut -c1-100
done
echo "budget 48, 64x704 and 1x352, 64x352"; PV2_PAR_CTA_BUDGET=48 timeout 200 python bench_head.py --batches 1,64 --sizes 352,704 --iters 10 2>&1 | grep -E "ms_graph" | cut -c1-100
echo "budget 96, 64x704 and 1x352, 64x352"; PV2_PAR_CTA_BUDGET=96 timeout 200 python bench_head.py --batches 1,64 --sizes 352,704 --iters 10 2>&1 | grep -E "ms_graph" | cut -c1-100
