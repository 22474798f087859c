#!/bin/bash
mkdir -p gpurun_out
for kv in PV2_STREAMS=0 PV2_PDL=0 PV2_CONV_V1=1 PV2_WGRAD_STREAMS=0 PV2_BN_BWD_LEAN=0 PV2_BN_BWD_FUSED=1 PV2_BIL_BWD2=0 PV2_UP2_TILED=0 PV2_CONV_NARROW=0 PV2_CHAIN_PRIORITY=0; do
  echo "== $kv"; env $kv timeout 600 python -m pytest tests/test_gpu_bench_config.py tests/test_gpu_models.py -x -q -m gpu --timeout 200 2>&1 | grep -E "^E   |passed|failed|^FAILED|Timeout" | head -5 | cut -c1-250
done
