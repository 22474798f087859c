"""The tcgen05 conv engine (BasicConv2d = conv + BN [+ ReLU], RFB, aggregation) against the CPU oracle:
forward, running-stat updates and every gradient (input, weight, gamma, beta), in both precisions.
fp32 mode (tf32 hi/lo split, 3 products) must hold the north_star's 1e-3; bf16 mode the stated 2e-2."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import pranet_v2_b200 as P
from pranet_v2_b200 import engine as E
from pranet_v2_b200.heads import BasicConv2d, RFB_modified, aggregation
from oracle import dsra_oracle as O
from oracle import synth, templates

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = {"fp32": 1e-3, "bf16": 2e-2}


@pytest.fixture(autouse=True)
def _restore_precision():
    yield
    E.set_precision("auto")


def _rel(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-6)


GEOMS = [  # cin, cout, kernel, padding, dilation, H, W, B
    (64, 64, 3, 1, 1, 22, 22, 2),
    (256, 256, 5, 2, 1, 11, 11, 2),
    (2048, 256, 1, 0, 1, 11, 11, 2),
    (512, 64, 1, 0, 1, 44, 44, 1),
    (32, 32, (1, 7), (0, 3), 1, 22, 22, 2),
    (32, 32, (5, 1), (2, 0), 1, 12, 12, 2),
    (32, 32, 3, 7, 7, 44, 44, 1),
    (128, 32, 3, 1, 1, 8, 8, 3),
    (96, 96, 3, 1, 1, 44, 44, 1),
    (64, 1, 3, 1, 1, 22, 22, 2),
    (256, 3, 1, 0, 1, 3, 3, 2),
    (320, 9, 3, 1, 1, 14, 14, 2),
    (64, 64, 3, 1, 1, 7, 5, 2),
    # persistent-kernel cases: more 128-pixel tiles than SMs (a CTA walks several work items through one ring and both TMEM
    # accumulators), and an N of two 208-column tiles
    (32, 32, 3, 1, 1, 88, 88, 4),
    (64, 64, 3, 1, 1, 88, 88, 4),
    (256, 416, 1, 0, 1, 11, 11, 16),
]


@pytest.mark.parametrize("geom", GEOMS)
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("training", [True, False])
def test_basic_conv2d(geom, precision, training):
    cin, cout, k, pad, dil, H, W, B = geom
    E.set_precision(precision)
    m = BasicConv2d(cin, cout, k, padding=pad, dilation=dil)
    sd = synth.synth_state_dict({"m." + kk: v for kk, v in m.state_dict().items()}, seed=21)
    if precision == "bf16":   # what the bf16 tensor-core path consumes; keeps the oracle's ReLU mask identical
        sd["m.conv.weight"] = sd["m.conv.weight"].bfloat16().float()
    m.load_state_dict({kk[2:]: v for kk, v in sd.items()})
    m = m.to(DEV).train(training)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, cin, H, W, generator=g)
    gout = torch.randn(B, cout, H, W, generator=g)
    if precision == "bf16":
        x = x.bfloat16().float()
    # oracle (fp32 CPU) with autograd
    ref_sd = {kk: v.clone() for kk, v in sd.items()}
    for kk in ("m.conv.weight", "m.bn.weight", "m.bn.bias"):
        ref_sd[kk].requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    ref = F.relu(O.basic_conv(xr, ref_sd, "m", training, pad, dil))
    ref.backward(gout)
    xd = x.to(DEV).requires_grad_(True)
    out = m(xd, relu=True)
    out.backward(gout.to(DEV))
    tol = TOL[precision]
    assert out.shape == ref.shape
    assert (out.cpu() - ref).abs().max().item() <= tol * max(1.0, ref.abs().max().item())
    if training:
        np.testing.assert_allclose(m.bn.running_mean.cpu().numpy(), ref_sd["m.bn.running_mean"].detach().numpy(), rtol=tol, atol=tol * 0.1)
        np.testing.assert_allclose(m.bn.running_var.cpu().numpy(), ref_sd["m.bn.running_var"].detach().numpy(), rtol=tol, atol=tol * 0.1)
        assert int(m.bn.num_batches_tracked) == 1
    gtol = 3 * tol
    assert _rel(xd.grad.cpu(), xr.grad) <= gtol, "dx"
    assert _rel(m.conv.weight.grad.cpu(), ref_sd["m.conv.weight"].grad) <= gtol, "dW"
    assert _rel(m.bn.weight.grad.cpu(), ref_sd["m.bn.weight"].grad) <= gtol, "dgamma"
    assert _rel(m.bn.bias.grad.cpu(), ref_sd["m.bn.bias"].grad) <= gtol, "dbeta"


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("cin,hw", [(512, 12), (2048, 5)])
def test_rfb(precision, cin, hw):
    E.set_precision(precision)
    m = RFB_modified(cin, 32)
    sd = synth.synth_state_dict(templates.rfb("r", cin, 32), seed=4)
    m.load_state_dict({k[2:]: v for k, v in sd.items()})
    m = m.to(DEV).train()
    x = torch.relu(torch.randn(2, cin, hw, hw, generator=torch.Generator().manual_seed(1)))
    if precision == "bf16":
        x = x.bfloat16().float()
    xr = x.clone().requires_grad_(True)
    ref = O.rfb(xr, {k: v.clone() for k, v in sd.items()}, "r", True)
    gout = torch.randn(ref.shape, generator=torch.Generator().manual_seed(2))
    ref.backward(gout)
    xd = x.to(DEV).requires_grad_(True)
    out = m(xd)
    out.backward(gout.to(DEV))
    tol = TOL[precision] * (1 if precision == "fp32" else 3)   # 5 stacked bf16 layers
    assert (out.cpu() - ref).abs().max().item() <= tol * max(1.0, ref.abs().max().item())
    assert _rel(xd.grad.cpu(), xr.grad) <= 5 * tol


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("num_class", [1, 3, None])
def test_aggregation(precision, num_class):
    E.set_precision(precision)
    m = aggregation(32, num_class)
    sd = synth.synth_state_dict(templates.aggregation("a", 32, num_class), seed=6)
    m.load_state_dict({k[2:]: v for k, v in sd.items()})
    m = m.to(DEV).train()
    g = torch.Generator().manual_seed(3)
    xs = [torch.relu(torch.randn(2, 32, s, s, generator=g)) for s in (3, 6, 12)]
    if precision == "bf16":
        xs = [x.bfloat16().float() for x in xs]
    xr = [x.clone().requires_grad_(True) for x in xs]
    osd = {k: v.clone() for k, v in sd.items()}
    ref = O.aggregation_v1(*xr, osd, "a", True) if num_class is None else O.aggregation_v2(*xr, osd, "a", True)
    ref = (ref,) if num_class is None else ref
    gouts = [torch.randn(r.shape, generator=g) for r in ref]
    torch.autograd.backward(ref, gouts)
    xd = [x.to(DEV).requires_grad_(True) for x in xs]
    out = m(*xd)
    out = (out,) if num_class is None else out
    torch.autograd.backward(out, [t.to(DEV) for t in gouts])
    tol = TOL[precision] * (1 if precision == "fp32" else 4)
    for o, r in zip(out, ref):
        assert (o.cpu() - r).abs().max().item() <= tol * max(1.0, r.abs().max().item())
    for a, b in zip(xd, xr):
        assert _rel(a.grad.cpu(), b.grad) <= 5 * tol


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("B,H,W", [(16, 44, 44), (5, 40, 40), (3, 11, 11)])
def test_fused_bn_statistics_many_tiles(precision, B, H, W):
    """BatchNorm statistics produced by the conv epilogue + ticket fold (one and two fold levels, ragged last tile) against
    torch's batch_norm on the same conv output: mean / biased var, running stats, and the BN backward sums."""
    E.set_precision(precision)
    m = BasicConv2d(64, 96, 3, padding=1)
    sd = synth.synth_state_dict({"m." + kk: v for kk, v in m.state_dict().items()}, seed=33)
    if precision == "bf16":
        sd["m.conv.weight"] = sd["m.conv.weight"].bfloat16().float()
    m.load_state_dict({kk[2:]: v for kk, v in sd.items()})
    m = m.to(DEV).train()
    x = torch.randn(B, 64, H, W, generator=torch.Generator().manual_seed(9))
    if precision == "bf16":
        x = x.bfloat16().float()
    xd = x.to(DEV).requires_grad_(True)
    y = m(xd, relu=True)
    gout = torch.randn(y.shape, generator=torch.Generator().manual_seed(10))
    y.backward(gout.to(DEV))
    xr = x.clone().requires_grad_(True)
    w = sd["m.conv.weight"].clone().requires_grad_(True)
    gam, bet = sd["m.bn.weight"].clone().requires_grad_(True), sd["m.bn.bias"].clone().requires_grad_(True)
    rm, rv = sd["m.bn.running_mean"].clone(), sd["m.bn.running_var"].clone()
    ref = F.relu(F.batch_norm(F.conv2d(xr, w, padding=1), rm, rv, gam, bet, True, 0.1, 1e-5))
    ref.backward(gout)
    tol = TOL[precision]
    assert _rel(y.detach().cpu(), ref.detach()) <= tol
    assert _rel(m.bn.running_mean.cpu(), rm) <= tol and _rel(m.bn.running_var.cpu(), rv) <= tol
    assert int(m.bn.num_batches_tracked) == int(sd["m.bn.num_batches_tracked"]) + 1
    assert _rel(xd.grad.cpu(), xr.grad) <= 5 * tol
    assert _rel(m.conv.weight.grad.cpu(), w.grad) <= 5 * tol
    assert _rel(m.bn.weight.grad.cpu(), gam.grad) <= 5 * tol and _rel(m.bn.bias.grad.cpu(), bet.grad) <= 5 * tol
