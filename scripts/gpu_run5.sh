#!/bin/bash
mkdir -p gpurun_out
for cfg in "dflt:" "s64:PV2_CONV_SMALL=64" "s128:PV2_CONV_SMALL=128" "s64c4:PV2_CONV_SMALL=64 PV2_CONV_CTAS=4" "all1deep:PV2_CONV_SMALL=100000"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 200 python bench_head.py --batches 16 --sizes 352 --iters 30 > gpurun_out/head_$name.log 2>&1
  echo "$name [$envs]: $(tail -1 gpurun_out/head_$name.log | cut -c1-120)"
done
