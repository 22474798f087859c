"""Parity of the configuration bench.py / bench_head.py actually time: the whole PraNet-V2 head at B = 16 x 352^2 in train
mode, replayed from a captured CUDA graph (twelve parallel branches on side streams, weight-gradient companion streams,
split-K layers, 242-tile persistent grids), forward AND backward, in bf16 (tolerance 2e-2, north_star) and fp32 (1e-3 / 2e-3).

The oracle (CPU fp32, oracle/dsra_oracle.py) is fed the same features and the same weights; in the bf16 case both are first
rounded to bf16 -- what the tensor-core path consumes -- so the comparison measures the kernels, not the input rounding.
Reference call sequence: binary_seg/MyTrain_med.py:76-86 (model forward, four structure losses, backward)."""
import pytest
import torch

import pranet_v2_b200 as P
from pranet_v2_b200 import engine as E
from oracle import dsra_oracle as O
from oracle import golden_cases as G
from oracle import synth

pytestmark = pytest.mark.gpu
DEV = "cuda"
B, S = 16, 352


@pytest.fixture(autouse=True)
def _restore_precision():
    yield
    E.set_precision("auto")


def _graph_step(m, feats, gt):
    """Head forward + 4x structure loss + backward captured in ONE CUDA graph (like bench_head.sweep_point / TrainStep) and
    replayed; returns (outs, loss, feature grads, parameter grads) of the replay."""
    params = m.head_parameters()
    state = {k: v.clone() for k, v in m.state_dict().items()}

    def step():
        for p in params:
            p.grad = None
        for f in feats:
            f.grad = None
        outs = m.forward_head(*feats)
        loss = P.structure_loss_multi([(outs[i], outs[i + 4]) for i in range(4)], gt).sum()
        loss.backward()
        return outs, loss

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        outs, loss = step()
    m.load_state_dict(state)           # the warm-up / capture runs updated the BatchNorm running statistics
    graph.replay()
    torch.cuda.synchronize()
    return outs, loss, [f.grad for f in feats], {k: p.grad for k, p in m.named_parameters() if p.grad is not None}, graph


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_head_b16_352_graph_fwd_bwd(precision):
    E.set_precision(precision)
    m = P.PraNet_V2(num_class=1)
    tmpl = {k: v for k, v in m.state_dict().items() if G.head_key_filter(k)}
    sd = synth.synth_state_dict(tmpl, seed=23)
    feats_cpu = synth.backbone_features(B, S, 23)
    if precision == "bf16":       # what the bf16 tensor-core path consumes
        sd = {k: (v.bfloat16().float() if (v.dtype == torch.float32 and k.endswith("conv.weight")) else v) for k, v in sd.items()}
        feats_cpu = [f.bfloat16().float() for f in feats_cpu]
    m.load_state_dict(sd, strict=False)
    m = m.to(DEV).train()
    gt = synth.ellipse_masks(B, S, S, 23)
    if precision == "bf16":
        feats = [f.to(DEV).bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True) for f in feats_cpu]
    else:
        feats = [f.to(DEV).requires_grad_(True) for f in feats_cpu]
    gt_dev = gt.to(DEV)          # must outlive the graph: a replay reads the mask from this very buffer
    outs, loss, dfeats, dparams, graph = _graph_step(m, feats, gt_dev)

    rfeats = [f.clone().requires_grad_(True) for f in feats_cpu]
    rsd = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in sd.items()}
    ref = O.pranet_v2_head(*rfeats, rsd, training=True)
    rloss = sum(O.structure_loss(ref[i], ref[i + 4], gt, 1 - gt) for i in range(4))
    rloss.backward()

    # fp32: the north_star's 1e-3 max-abs.  bf16: 2e-2 of the logit scale -- every activation between the ~25 stacked conv + BN
    # layers is rounded to bf16 (relative step 3.9e-3) on our side and kept in fp32 by the oracle, and the final maps carry the
    # factor 2 of the C = 1 fusion (fg + fg * softmax_1 = 2 fg), so the error is proportional to the logits' magnitude
    errs = [(o.float().cpu() - r.detach()).abs().max().item() for o, r in zip(outs, ref)]
    mags = [r.detach().abs().max().item() for r in ref]
    report = ", ".join(f"out{i}: {e:.2e} (|ref| <= {mg:.1f})" for i, (e, mg) in enumerate(zip(errs, mags)))
    print(f"[{precision}] logits max-abs error: {report}")
    for i, (e, mg) in enumerate(zip(errs, mags)):
        tol = 2e-2 * max(1.0, mg) if precision == "bf16" else 1e-3        # bf16: the per-layer tests' metric, max|err| / max|ref| <= 2e-2
        assert e <= tol, f"{precision} out{i}: max-abs {e:.3e} > {tol:.3e}; all: {report}"
    if precision == "bf16":
        assert max(errs) <= 6e-2, report                                      # and an absolute ceiling: 3 x the north_star's 2e-2 on |logits| <= ~8
    # prediction rule of MyTest_med.py:36-38: sigmoid(sum of the fg maps) > 0.5  <=>  sum > 0
    osum, rsum = sum(o.float().cpu() for o in outs[:4]), sum(r.detach() for r in ref[:4])
    same = (osum > 0) == (rsum > 0)
    agree = same.float().mean().item()
    if precision == "fp32":
        assert agree >= 0.999, f"mask agreement {agree:.5f}"
    else:
        # random-init logits are not bimodal: ~1 % of the pixels have a reference sum within the bf16 error of the threshold and may
        # legitimately fall on either side.  Every pixel whose reference sum clears the threshold by more than the summed bf16
        # tolerance of the four maps must agree (>= 99.9 %), and the overall agreement is reported and bounded too.
        margin = 4 * 2.5e-2
        clear = rsum.abs() > margin
        agree_clear = same[clear].float().mean().item()
        print(f"[bf16] mask agreement: all pixels {agree:.5f}, pixels with |reference sum| > {margin}: {agree_clear:.6f} ({clear.float().mean().item():.3f} of all)")
        assert agree_clear >= 0.999 and agree >= 0.995, (agree, agree_clear)
    lrel = abs(loss.item() - rloss.item()) / abs(rloss.item())
    assert lrel <= (2e-3 if precision == "bf16" else 1e-4), f"loss rel {lrel:.3e}"
    gtol = 5e-2 if precision == "bf16" else 2e-3
    for i, (g, rf) in enumerate(zip(dfeats, rfeats)):
        diff = g.float().cpu() - rf.grad
        rel_max = diff.abs().max().item() / rf.grad.abs().max().item()
        rel_l2 = diff.double().norm().item() / rf.grad.double().norm().item()
        print(f"[{precision}] dfeat{i}: max-abs / max {rel_max:.3e}, L2 relative {rel_l2:.3e}")
        if precision == "fp32":
            # as a vector the gradient holds 2e-3; single elements may be off by more where a ReLU input is within rounding of zero
            # (the mask bit then differs between the tf32x3 tensor-core sum and ATen's and re-routes one path)
            assert rel_l2 <= gtol and rel_max <= 2e-2, f"fp32 dfeat{i}: L2 relative {rel_l2:.3e}, max {rel_max:.3e}"
        else:
            # bf16: gradients pass ~25 layers as bf16 operands and through ReLU masks of bf16-rounded activations, so single
            # elements can be far off (a mask bit that differs re-routes a whole path); the gradient as a vector must still agree
            assert rel_l2 <= 8e-2 and rel_max <= 0.3, f"bf16 dfeat{i}: L2 relative {rel_l2:.3e}, max {rel_max:.3e}"
    checked = 0
    for k, g in dparams.items():
        r = rsd.get(k)
        if r is None or r.grad is None:
            continue
        rn = r.grad.double().norm().item()
        gn = g.double().norm().item()
        assert abs(gn - rn) <= gtol * rn + 1e-7, f"{precision} grad norm of {k}: {gn:.6e} vs {rn:.6e}"
        if g.numel() >= 1024:     # direction too, for the weight tensors
            cos = torch.nn.functional.cosine_similarity(g.double().flatten().cpu(), r.grad.double().flatten(), dim=0).item()
            assert cos >= (0.995 if precision == "bf16" else 0.99999), f"{precision} grad direction of {k}: cos {cos:.6f}"
        checked += 1
    assert checked >= 150, checked
    # the BatchNorm running statistics after exactly one replayed step
    post = m.state_dict()
    nstat = 0
    for k, v in rsd.items():       # the oracle updated its copy in place, like nn.BatchNorm2d
        if k.endswith(("running_mean", "running_var")):
            torch.testing.assert_close(post[k].float().cpu(), v.detach(), rtol=2e-2 if precision == "bf16" else 1e-3, atol=1e-3 if precision == "bf16" else 1e-5)
            nstat += 1
        elif k.endswith("num_batches_tracked"):
            assert int(post[k]) == int(v)
    assert nstat >= 100, nstat
    # a second replay from the same inputs reproduces the first (split-K reductions aside, to rounding)
    l1 = loss.item()
    graph.replay()
    torch.cuda.synchronize()
    assert abs(loss.item() - l1) <= 1e-5 * abs(l1)
