#!/bin/bash
# conv engine bring-up: a tiny direct probe first (bounded), then the test files
mkdir -p gpurun_out
timeout 120 python - > gpurun_out/probe.log 2>&1 <<'PY'
import torch, time
import pranet_v2_b200 as P
from pranet_v2_b200 import engine as E
from pranet_v2_b200.heads import BasicConv2d
import torch.nn.functional as F
torch.manual_seed(0)
for prec in ("bf16", "fp32"):
    E.set_precision(prec)
    for (cin, cout, k, pad, H) in ((64, 64, 1, 0, 16), (64, 64, 3, 1, 16), (256, 256, 5, 2, 11), (512, 32, 1, 0, 44)):
        m = BasicConv2d(cin, cout, k, padding=pad).cuda().eval()
        x = torch.randn(2, cin, H, H, device="cuda")
        if prec == "bf16":
            x = x.bfloat16().float()
            w = m.conv.weight.data.bfloat16().float()
        else:
            w = m.conv.weight.data
        with torch.no_grad():
            out = m(x)
            ref = F.batch_norm(F.conv2d(x.double(), w.double(), None, 1, pad), m.bn.running_mean.double(), m.bn.running_var.double(), m.bn.weight.double(), m.bn.bias.double(), False, 0.1, 1e-5).float()
        torch.cuda.synchronize()
        print(prec, cin, cout, k, H, "max-abs err", (out - ref).abs().max().item(), "ref max", ref.abs().max().item(), flush=True)
PY
echo "probe exit $?" >> gpurun_out/probe.log
tail -12 gpurun_out/probe.log
timeout 900 python -m pytest tests/test_gpu_conv.py -q --timeout 120 -x 2>&1 | tail -30 > gpurun_out/pytest_conv.log
tail -15 gpurun_out/pytest_conv.log
