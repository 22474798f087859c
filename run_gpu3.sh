#!/bin/bash
mkdir -p gpurun_out
timeout 300 python - > gpurun_out/probe2.log 2>&1 <<'PY'
import torch, torch.nn.functional as F
import pranet_v2_b200 as P
from pranet_v2_b200 import engine as E
from pranet_v2_b200.heads import BasicConv2d
from oracle import dsra_oracle as O, synth
def rel(a,b): return ((a-b).abs().max()/b.abs().max().clamp_min(1e-9)).item()
for prec in ("bf16","fp32"):
  for training in (False, True):
    for relu in (False, True):
      for (cin,cout,k,pad,H,B) in ((64,64,1,0,16,1),(64,64,3,1,22,2),(256,256,5,2,11,2)):
        E.set_precision(prec)
        m = BasicConv2d(cin,cout,k,padding=pad)
        sd = synth.synth_state_dict({"m."+kk:v for kk,v in m.state_dict().items()}, seed=21)
        m.load_state_dict({kk[2:]:v for kk,v in sd.items()}); m = m.cuda().train(training)
        g = torch.Generator().manual_seed(5)
        x = torch.randn(B,cin,H,H,generator=g); gout = torch.randn(B,cout,H,H,generator=g)
        if prec=="bf16": x = x.bfloat16().float()
        rsd = {kk:v.clone() for kk,v in sd.items()}
        for kk in ("m.conv.weight","m.bn.weight","m.bn.bias"): rsd[kk].requires_grad_(True)
        xr = x.clone().requires_grad_(True)
        ref = O.basic_conv(xr, rsd, "m", training, pad, 1)
        if relu: ref = F.relu(ref)
        ref.backward(gout)
        xd = x.cuda().requires_grad_(True)
        out = m(xd, relu=relu); out.backward(gout.cuda())
        print(prec, "train" if training else "eval", "relu" if relu else "lin", (cin,cout,k,H,B),
              "out %.1e dx %.1e dW %.1e dg %.1e db %.1e" % (rel(out.cpu(),ref), rel(xd.grad.cpu(), xr.grad), rel(m.conv.weight.grad.cpu(), rsd["m.conv.weight"].grad),
               rel(m.bn.weight.grad.cpu(), rsd["m.bn.weight"].grad), rel(m.bn.bias.grad.cpu(), rsd["m.bn.bias"].grad)), flush=True)
PY
echo "exit $?" >> gpurun_out/probe2.log
cat gpurun_out/probe2.log | tail -30
