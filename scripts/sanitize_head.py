"""One small head step (forward + 4x structure loss + backward) for compute-sanitizer: every pv2 kernel of the training path runs
once, in fp32 (tf32x3) and bf16, eager (no graph), side streams on.  Usage on the GPU box:
    compute-sanitizer --tool memcheck  python scripts/sanitize_head.py
    compute-sanitizer --tool racecheck python scripts/sanitize_head.py
    compute-sanitizer --tool initcheck python scripts/sanitize_head.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import pranet_v2_b200 as P
from pranet_v2_b200 import synthetic

B, S = int(os.environ.get("SAN_B", "2")), int(os.environ.get("SAN_S", "96"))
dev = "cuda"
torch.manual_seed(0)
for prec in os.environ.get("SAN_PREC", "bf16,fp32").split(","):
    P.set_precision(prec)
    model = P.PraNet_V2(num_class=1).to(dev).train()
    g = torch.Generator().manual_seed(1)
    feats = [torch.relu(torch.randn(B, c, S // s, S // s, generator=g)).to(dev) for c, s in ((512, 8), (1024, 16), (2048, 32))]
    if prec == "bf16":
        feats = [f.bfloat16().contiguous(memory_format=torch.channels_last) for f in feats]
    feats = [f.requires_grad_(True) for f in feats]
    gt = synthetic.ellipse_masks(B, S, S, 3).to(dev)
    for it in range(2):
        outs = model.forward_head(*feats)
        loss = P.structure_loss_multi([(outs[i], outs[i + 4]) for i in range(4)], gt, prepared=P.ops.structure_loss_prepare(gt)).sum()
        loss.backward()
        torch.cuda.synchronize()
        print(prec, it, "loss", float(loss), "finite grads", all(torch.isfinite(f.grad.float()).all().item() for f in feats))
print("done")
