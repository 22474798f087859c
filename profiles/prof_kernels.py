"""Small driver for ncu captures of the pv2 kernels (not a benchmark: numbers printed under a profiler are never
bench values).  Usage (on the GPU box, one GPU):
    ncu --set full --clock-control none --import-source on -k regex:structure_loss -c 6 -o gpurun_out/prof_loss  python profiles/prof_kernels.py loss
    ncu --set full --clock-control none --import-source on -k regex:conv_ -c 120 -o gpurun_out/prof_conv         python profiles/prof_kernels.py head
    ncu --set full --clock-control none --import-source on -k regex:adam_clamp -c 3 -o gpurun_out/prof_adam       python profiles/prof_kernels.py adam
    ncu --set full --clock-control none --import-source on -k regex:tail_ -c 4 -o gpurun_out/prof_tail            python profiles/prof_kernels.py tail
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pranet_v2_b200 as P
from pranet_v2_b200 import synthetic

what = sys.argv[1] if len(sys.argv) > 1 else "loss"
B, S = (int(v) for v in os.environ.get("PV2_PROF_BS", "16x352").lower().split("x"))
dev = "cuda"
torch.manual_seed(0)
if what == "step":     # two eager training steps of the bench workload (launch list of the pv2 kernels)
    from pranet_v2_b200.train import TrainStep
    ts = TrainStep(P.PraNet_V2(num_class=1), device="cuda:0", use_graph=False)
    x, gt = synthetic.images(B, S, 1).to(dev), synthetic.ellipse_masks(B, S, S, 1).to(dev)
    for it in range(2):
        ts.step_device(x, gt)
elif what == "adam":   # the optimizer tail over the PraNet-V2 Res2Net-50 parameter count
    lib = P._lib.load()
    n = 30_499_908
    p, g = torch.randn(n, device=dev), torch.randn(n, device=dev)
    m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    step, ticket = torch.zeros(1, dtype=torch.int64, device=dev), torch.zeros(1, dtype=torch.int32, device=dev)
    for it in range(3):
        P._lib.check(lib.pv2_adam_clamp_flat(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, step.data_ptr(), ticket.data_ptr(),
                                             1e-4, 0.9, 0.999, 1e-8, 0.0, 0, 0.5, 1.0, torch.cuda.current_stream().cuda_stream), "adam")
elif what == "tail":   # fused inference tails from the low-res maps
    maps = [torch.randn(B, 1, S // s, S // s, device=dev) for s in (8, 16, 32, 8)]
    for it in range(2):
        P.ops.infer_tail_binary(maps, [8, 16, 32, 8], (S, S))
    fg = [torch.randn(B, 9, 224 // s, 224 // s, device=dev) for s in (32, 16, 8, 4)]
    bg = [torch.randn(B, 9, 224 // s, 224 // s, device=dev) for s in (32, 16, 8, 4)]
    P.ops.infer_tail_argmax(fg, bg, [32, 16, 8, 4])
elif what == "lowres":  # loss from the low-res maps next to the four launches it replaces; 7 pv2 launches per iteration:
    # lowres fwd, lowres bwd, fold | bilinear x8 fwd, loss fwd, loss bwd, bilinear x8 bwd
    import ctypes
    from pranet_v2_b200.ops import PV2_F32, _ratio
    lib = P._lib.load()
    scs = (8, 16, 32, 8)
    m = synthetic.ellipse_masks(B, S, S, 3).to(dev)
    ihs = (ctypes.c_int * 8)(*[S // s for s in scs * 2])
    rr = (ctypes.c_float * 8)(*[_ratio(S // s, S, False, float(s)) for s in scs * 2])
    st = lambda: torch.cuda.current_stream().cuda_stream
    for it in range(2):
        pairs = [(torch.randn(B, 1, S // s, S // s, device=dev).requires_grad_(True), torch.randn(B, 1, S // s, S // s, device=dev).requires_grad_(True))
                 for s in scs]
        P.structure_loss_lowres(pairs, list(scs), m).sum().backward()
        lows = [a.detach() for a, _ in pairs] + [b.detach() for _, b in pairs]
        his = [torch.empty(B, 1, S, S, device=dev) for _ in range(8)]
        (pl, k1), (ph, k2) = P._lib.ptr_array(lows), P._lib.ptr_array(his)
        P._lib.check(lib.pv2_bilinear_multi_fwd(pl, ph, ihs, ihs, rr, rr, 8, B, S, S, 0, PV2_F32, st()), "bilinear_multi_fwd")
        ups = [t.requires_grad_(True) for t in his]
        P.structure_loss_multi([(ups[i], ups[i + 4]) for i in range(4)], m).sum().backward()
        gs, dl = [t.grad for t in ups], [torch.empty_like(t) for t in lows]
        (pg, k3), (pd, k4) = P._lib.ptr_array(gs), P._lib.ptr_array(dl)
        P._lib.check(lib.pv2_bilinear_multi_bwd(pg, pd, ihs, ihs, rr, rr, 8, B, S, S, 0, PV2_F32, st()), "bilinear_multi_bwd")
        err = max((d - t.grad).abs().max().item() / t.grad.abs().max().item() for d, t in zip(dl, [a for a, _ in pairs] + [b for _, b in pairs]))
        print("fused vs unfused low-res gradient, max rel err", err)
elif what == "r2":     # round 2: the kernels the bench line's roofline / other_kernels name, at the benchmarked sizes
    import ctypes
    import torch.nn as nn
    from pranet_v2_b200.ops import PV2_F32, _ratio
    lib = P._lib.load()
    m = synthetic.ellipse_masks(B, S, S, 3).to(dev)
    for it in range(2):
        pairs = [(torch.randn(B, 1, S, S, device=dev).requires_grad_(True), torch.randn(B, 1, S, S, device=dev).requires_grad_(True)) for _ in range(4)]
        P.structure_loss_multi(pairs, m, prepared=P.ops.structure_loss_prepare(m)).sum().backward()      # boundary_weight, loss fwd (streaming), loss bwd
        P.structure_loss_multi([(a.detach(), b.detach()) for a, b in pairs], m)                            # fused forward
        scs = (8, 16, 32, 8) * 2
        lows = [torch.randn(B, 1, S // s, S // s, device=dev) for s in scs]
        his = [torch.randn(B, 1, S, S, device=dev) for _ in scs]
        ihs = (ctypes.c_int * 8)(*[S // s for s in scs])
        rr = (ctypes.c_float * 8)(*[_ratio(S // s, S, False, float(s)) for s in scs])
        (pl, k1), (ph, k2) = P._lib.ptr_array(lows), P._lib.ptr_array(his)
        st = torch.cuda.current_stream().cuda_stream
        P._lib.check(lib.pv2_bilinear_multi_fwd(pl, ph, ihs, ihs, rr, rr, 8, B, S, S, 0, PV2_F32, st), "bilinear_multi_fwd")
        P._lib.check(lib.pv2_bilinear_multi_bwd(ph, pl, ihs, ihs, rr, rr, 8, B, S, S, 0, PV2_F32, st), "bilinear_multi_bwd")
        fg = [torch.randn(B, 9, 224, 224, device=dev).requires_grad_(True) for _ in range(8)]
        lab = torch.randint(0, 9, (B, 224, 224), device=dev)
        P.mc_dual_loss(fg[:4], fg[4:], lab, 9).backward()
        eng = P.engine.Engine(torch.device(dev), "bf16", True, False)
        for (cin, cout, k, hw) in ((512, 224, 1, S // 8), (1024, 224, 1, S // 16), (2048, 416, 1, S // 32), (256, 256, 5, S // 32), (96, 96, 3, S // 8), (64, 64, 3, S // 8), (32, 32, 3, S // 8)):
            conv = nn.Conv2d(cin, cout, k, padding=k // 2, bias=False).to(dev)
            bnm = nn.BatchNorm2d(cout).to(dev).train()
            a = eng.new_act(B, hw, hw, cin)
            a.t.normal_()
            eng.conv(a, [conv], [bnm])
elif what == "r2conv":  # the seven distinct forward conv shapes of the head with the BatchNorm statistics fused, at PV2_PROF_BS
    import torch.nn as nn
    eng = P.engine.Engine(torch.device(dev), "bf16", True, False)
    for (cin, cout, k, hw) in ((512, 224, 1, S // 8), (1024, 224, 1, S // 16), (2048, 416, 1, S // 32), (256, 256, 5, S // 32), (96, 96, 3, S // 8), (64, 64, 3, S // 8)):
        conv = nn.Conv2d(cin, cout, k, padding=k // 2, bias=False).to(dev)
        bnm = nn.BatchNorm2d(cout).to(dev).train()
        a = eng.new_act(B, hw, hw, cin)
        a.t.normal_()
        for it in range(2):
            eng.conv(a, [conv], [bnm])
elif what == "loss":
    m = synthetic.ellipse_masks(B, S, S, 3).to(dev)
    for it in range(2):
        pairs = [(torch.randn(B, 1, S, S, device=dev).requires_grad_(True), torch.randn(B, 1, S, S, device=dev).requires_grad_(True)) for _ in range(4)]
        P.structure_loss_multi(pairs, m).sum().backward()
else:
    model = P.PraNet_V2(num_class=1).to(dev).train()
    feats = [torch.relu(torch.randn(B, c, S // s, S // s, device=dev)).bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
             for c, s in ((512, 8), (1024, 16), (2048, 32))]
    gt = synthetic.ellipse_masks(B, S, S, 3).to(dev)
    for it in range(2):
        outs = model.forward_head(*feats)
        P.structure_loss_multi([(outs[i], outs[i + 4]) for i in range(4)], gt).sum().backward()
torch.cuda.synchronize()
print("done", what)
