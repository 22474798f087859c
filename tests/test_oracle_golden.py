"""The CPU oracle against the golden vectors frozen from the UNMODIFIED reference
(oracle/make_golden.py).  This is what pins the oracle (the reference has no
tests of its own for this path)."""
import numpy as np
import pytest
import torch

from oracle import dsra_oracle as O
from oracle import golden_cases as G
from oracle import synth, templates

torch.set_num_threads(4)


def _sub(t, stride):
    t = t.detach()
    return (t[:, :, ::stride, ::stride] if stride > 1 else t).numpy()


@pytest.mark.parametrize("name", list(G.LOWRES_LOSS_CASES))
def test_structure_loss_lowres_oracle(name):
    """Oracle of the loss from the low-res maps == the reference's own statements (pranet.py final F.interpolate calls +
    MyTrain_med.py:74,78-82 compiled from source), losses and gradients w.r.t. the eight low-res maps."""
    g = G.load(name)
    maps, m = G.lowres_loss_inputs(name)
    maps = [t.requires_grad_(True) for t in maps]
    losses = O.structure_loss_lowres([(maps[i], maps[i + 4]) for i in range(4)], G.lowres_loss_scales(name), m)
    losses.sum().backward()
    np.testing.assert_allclose(losses.detach().numpy(), g["losses"], rtol=1e-6)
    assert abs(losses.sum().item() - float(g["loss"])) <= 1e-6 * abs(float(g["loss"]))
    for i, t in enumerate(maps):
        np.testing.assert_allclose(t.grad.numpy(), g[f"d{i}"], rtol=1e-5, atol=1e-7 * np.abs(g[f"d{i}"]).max())


@pytest.mark.parametrize("name", list(G.STRUCTURE_LOSS_CASES))
def test_structure_loss_oracle(name):
    g = G.load(name)
    stride = G.STRUCTURE_LOSS_CASES[name][5]
    pred, pred_bg, m, mb = G.structure_loss_inputs(name)
    assert abs(m.double().sum().item() - float(g["mask_sum"])) < 1e-6 * max(1.0, float(g["mask_sum"]))
    pred.requires_grad_(True)
    pred_bg.requires_grad_(True)
    loss = O.structure_loss(pred, pred_bg, m, mb)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= 1e-6 * abs(float(g["loss"]))
    np.testing.assert_allclose(_sub(pred.grad, stride), g["dpred"], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(_sub(pred_bg.grad, stride), g["dpred_bg"], rtol=1e-5, atol=1e-9)
    # the explicit float64 numpy restatement agrees with the reference as well
    l64, d64, db64 = O.structure_loss_np(pred.detach().numpy(), pred_bg.detach().numpy(), m.numpy(), mb.numpy())
    assert abs(l64 - float(g["loss"])) <= 2e-6 * abs(float(g["loss"]))
    scale = np.abs(g["dpred"]).max()
    np.testing.assert_allclose(d64[:, :, ::stride, ::stride], g["dpred"], rtol=0, atol=2e-5 * scale)
    np.testing.assert_allclose(db64[:, :, ::stride, ::stride], g["dpred_bg"], rtol=0, atol=2e-5 * np.abs(g["dpred_bg"]).max())


def _head_sd(case):
    v1 = case["model"] in ("PraNet", "PVT_PraNet")
    tmpl = templates.pranet_head(case["ch"], 32, case["kw"].get("num_class", 1), v1=v1)
    return synth.synth_state_dict(tmpl, seed=1)


@pytest.mark.parametrize("name", list(G.HEAD_CASES))
def test_head_oracle(name):
    case = G.HEAD_CASES[name]
    g = G.load(name)
    sd = _head_sd(case)
    feats = G.head_inputs(name)
    want_grad = name in G.HEAD_GRAD_CASES
    if want_grad:
        feats = [f.requires_grad_(True) for f in feats]
    ctx = torch.enable_grad() if want_grad else torch.no_grad()
    with ctx:
        if case["model"] in ("PraNet", "PVT_PraNet"):
            outs = O.pranet_v1_head(*feats, sd, training=case["training"])
        else:
            outs = O.pranet_v2_head(*feats, sd, use_softmax=case["kw"].get("use_softmax", True),
                                    sem_downsample=case["kw"].get("sem_downsample", 1), training=case["training"])
    assert tuple(outs[0].shape) == tuple(g["out_shape"])
    for i, o in enumerate(outs):
        np.testing.assert_allclose(_sub(o, case["stride"]), g[f"out{i}"], rtol=1e-4, atol=1e-5)
    for k in [k for k in g if k.startswith("stat:")]:
        np.testing.assert_allclose(sd[k[5:]].numpy(), g[k], rtol=1e-5, atol=1e-6)
    if want_grad:
        S = outs[0].shape[-1]
        gt = synth.ellipse_masks(case["B"], S, S, seed=7)
        loss = sum(O.structure_loss(outs[i], outs[i + 4], gt, 1 - gt) for i in range(4))
        loss.backward()
        assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
        for i, f in enumerate(feats):
            ref = g[f"dfeat{i}"]
            np.testing.assert_allclose(f.grad.numpy(), ref, rtol=0, atol=2e-4 * np.abs(ref).max())


@pytest.mark.parametrize("name", list(G.MC_CASES))
def test_multiclass_heads_oracle(name):
    case = G.MC_CASES[name]
    g = G.load(name)
    bn = case["kind"] != "mist"
    names = ("ConvBlock4", "ConvBlock3", "ConvBlock2", "ConvBlock1") if bn else ("out_head1", "out_head2", "out_head3", "out_head4")
    ks = (1, 3, 3, 3) if bn else (1, 1, 1, 1)
    sd = synth.synth_state_dict(templates.dual_heads(case["channels"], case["num_class"], names, ks, bn), seed=2)
    feats = [torch.from_numpy(g[f"d{i}"]) for i in range(4)]
    with torch.no_grad():
        outs = O.dual_heads_cascade(feats, sd, ks, case.get("use_softmax", True), case["training"], names, bn)
        ups = O.final_upsample(outs)
    for i, o in enumerate(outs):
        np.testing.assert_allclose(o.numpy(), g[f"out{i}"], rtol=1e-4, atol=1e-5)
    for i, o in enumerate(ups):
        np.testing.assert_allclose(_sub(o, 2), g[f"up{i}"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("name", list(G.MC_LOSS_CASES))
def test_mc_dual_loss_oracle(name):
    case = G.MC_LOSS_CASES[name]
    g = G.load(name)
    P_fg, P_bg, labels = G.mc_loss_inputs(name)
    for t in P_fg + P_bg:
        t.requires_grad_(True)
    subsets = O.powerset_subsets(4)
    assert [sum(1 << i for i in s) for s in subsets] == list(g["subsets"])
    loss = O.mc_dual_loss(P_fg, P_bg, labels, case["num_class"])
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    for i in range(4):
        np.testing.assert_allclose(P_fg[i].grad.numpy(), g[f"dfg{i}"], rtol=1e-4, atol=1e-8)
        np.testing.assert_allclose(P_bg[i].grad.numpy(), g[f"dbg{i}"], rtol=1e-4, atol=1e-8)


def test_bilinear_np_matches_aten():
    import torch.nn.functional as F
    x = torch.randn(2, 3, 11, 7, generator=torch.Generator().manual_seed(0))
    for s in (2, 4, 8, 32):
        ref = F.interpolate(x, scale_factor=s, mode="bilinear").numpy()
        np.testing.assert_allclose(O.bilinear_np(x.numpy(), 11 * s, 7 * s, False, s), ref, rtol=1e-5, atol=1e-6)
    ref = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True).numpy()
    np.testing.assert_allclose(O.bilinear_np(x.numpy(), 22, 14, True), ref, rtol=1e-5, atol=1e-6)
    x = torch.randn(1, 2, 44, 44, generator=torch.Generator().manual_seed(1))
    ref = F.interpolate(x, scale_factor=0.25, mode="bilinear").numpy()
    np.testing.assert_allclose(O.bilinear_np(x.numpy(), 11, 11, False, 0.25), ref, rtol=1e-5, atol=1e-6)
    ref = F.interpolate(x, size=(50, 61), mode="bilinear").numpy()
    np.testing.assert_allclose(O.bilinear_np(x.numpy(), 50, 61, False), ref, rtol=1e-4, atol=5e-5)  # ATen rounds the ratio in fp32
