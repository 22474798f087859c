#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_multiclass.py -x -q -m gpu --timeout 200 -k "mc or multiclass or emcad or merit or mist or dual" 2>&1 | grep -E "^E   |passed|failed|^FAILED|Timeout" | head -8 | cut -c1-250
timeout 300 python - <<'PY'
import torch, sys
sys.path.insert(0, '.')
import bench, pranet_v2_b200 as P
r = bench.roofline_mc_loss(P, torch.device('cuda:0'))
print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items() if k in ('achieved', 'avg_ms')}, {k: round(v, 3) for k, v in r['fwd'].items() if isinstance(v, float)})
PY
