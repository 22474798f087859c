"""Parity of the memory-bound kernels (through the C ABI) against the CPU oracle and the golden
vectors frozen from the reference.  Tolerances: fp32 loss 1e-4 relative (north_star), gradients
1e-3 of their max-abs; bf16 logits 2e-2."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import pranet_v2_b200 as P
from oracle import dsra_oracle as O
from oracle import golden_cases as G
from oracle import synth

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _sub(t, stride):
    t = t.detach().float().cpu()
    return (t[:, :, ::stride, ::stride] if stride > 1 else t).numpy()


@pytest.mark.parametrize("name", list(G.STRUCTURE_LOSS_CASES))
@pytest.mark.parametrize("pass_mask_bg", [False, True])
def test_structure_loss_golden(name, pass_mask_bg):
    g = G.load(name)
    stride = G.STRUCTURE_LOSS_CASES[name][5]
    pred, pred_bg, m, mb = [t.to(DEV) for t in G.structure_loss_inputs(name)]
    pred.requires_grad_(True)
    pred_bg.requires_grad_(True)
    loss = P.structure_loss(pred, pred_bg, m, mb if pass_mask_bg else None)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= 1e-4 * abs(float(g["loss"]))
    for got, ref in ((pred.grad, g["dpred"]), (pred_bg.grad, g["dpred_bg"])):
        np.testing.assert_allclose(_sub(got, stride), ref, rtol=0, atol=1e-3 * np.abs(ref).max() + 1e-12)


@pytest.mark.parametrize("shape,kind", [((16, 1, 352, 352), "hard"), ((4, 1, 256, 256), "soft"), ((2, 1, 448, 448), "soft"),
                                         ((3, 2, 100, 75), "hard"), ((1, 1, 704, 704), "hard")])
def test_structure_loss_vs_oracle_full_size(shape, kind):
    B, C, H, W = shape
    pred, pred_bg = synth.logits(shape, 5, "p"), synth.logits(shape, 5, "q")
    m = (synth.ellipse_masks(B * C, H, W, 5) if kind == "hard" else synth.soft_masks(B * C, H, W, 5)).view(shape)
    ref_p, ref_q = pred.clone().requires_grad_(True), pred_bg.clone().requires_grad_(True)
    ref = O.structure_loss(ref_p, ref_q, m, 1 - m)
    ref.backward()
    p, q = pred.to(DEV).requires_grad_(True), pred_bg.to(DEV).requires_grad_(True)
    loss = P.structure_loss(p, q, m.to(DEV), None)
    loss.backward()
    assert abs(loss.item() - ref.item()) <= 1e-4 * abs(ref.item())
    for got, want in ((p.grad, ref_p.grad), (q.grad, ref_q.grad)):
        assert (got.cpu() - want).abs().max().item() <= 1e-3 * want.abs().max().item()


def test_structure_loss_multi_and_bf16():
    shape = (4, 1, 352, 352)
    m = synth.ellipse_masks(4, 352, 352, 9).view(shape)
    pairs = [(synth.logits(shape, 9, f"p{k}"), synth.logits(shape, 9, f"q{k}")) for k in range(4)]
    ref = torch.stack([O.structure_loss(p, q, m, 1 - m) for p, q in pairs])
    dev_pairs = [(p.to(DEV).requires_grad_(True), q.to(DEV).requires_grad_(True)) for p, q in pairs]
    losses = P.structure_loss_multi(dev_pairs, m.to(DEV))
    assert torch.allclose(losses.cpu(), ref, rtol=1e-4, atol=0)
    # x4 launch == 4 single launches, bit for bit (same tiles, same reduction order)
    singles = torch.stack([P.structure_loss(p, q, m.to(DEV)) for p, q in dev_pairs])
    assert torch.equal(singles, losses)
    w = torch.tensor([1.0, 0.5, 2.0, 1.5], device=DEV)
    (losses * w).sum().backward()
    rp = [(p.clone().requires_grad_(True), q.clone().requires_grad_(True)) for p, q in pairs]
    (torch.stack([O.structure_loss(p, q, m, 1 - m) for p, q in rp]) * w.cpu()).sum().backward()
    for (p, q), (a, b) in zip(dev_pairs, rp):
        assert (p.grad.cpu() - a.grad).abs().max() <= 1e-3 * a.grad.abs().max()
        assert (q.grad.cpu() - b.grad).abs().max() <= 1e-3 * b.grad.abs().max()
    # bf16 logits: compare against the oracle evaluated on the bf16-rounded logits (tolerance 2e-2 stated)
    bp = [(p.bfloat16(), q.bfloat16()) for p, q in pairs]
    refb = torch.stack([O.structure_loss(p.float(), q.float(), m, 1 - m) for p, q in bp])
    db = [(p.to(DEV).requires_grad_(True), q.to(DEV).requires_grad_(True)) for p, q in bp]
    lb = P.structure_loss_multi(db, m.to(DEV))
    assert torch.allclose(lb.cpu(), refb, rtol=1e-4, atol=0)
    lb.sum().backward()
    assert db[0][0].grad.dtype == torch.bfloat16
    rr = bp[0][0].float().requires_grad_(True)
    O.structure_loss(rr, bp[0][1].float(), m, 1 - m).backward()
    assert (db[0][0].grad.float().cpu() - rr.grad).abs().max() <= 2e-2 * rr.grad.abs().max()


def test_structure_loss_errors():
    x = torch.zeros(1, 1, 8, 8, device=DEV)
    with pytest.raises(ValueError):
        P.structure_loss(x, x, torch.zeros(1, 1, 8, 9, device=DEV))
    with pytest.raises(TypeError):
        P.structure_loss(x.half(), x.half(), x)


@pytest.mark.parametrize("shape,kw", [
    ((2, 1, 44, 44), dict(scale_factor=8)), ((2, 1, 11, 11), dict(scale_factor=32)), ((2, 3, 22, 22), dict(scale_factor=16)),
    ((2, 1, 44, 44), dict(scale_factor=0.25)), ((2, 1, 11, 11), dict(scale_factor=2)), ((3, 32, 11, 11), dict(scale_factor=2, align_corners=True)),
    ((2, 9, 7, 7), dict(size=(14, 14))), ((1, 4, 13, 9), dict(size=(31, 20))), ((2, 9, 56, 56), dict(scale_factor=4)),
    ((1, 2, 5, 7), dict(scale_factor=3)), ((1, 1, 1, 1), dict(scale_factor=4)), ((1, 1, 64, 64), dict(scale_factor=4.0)),
    # the separable integer-scale backward (s % 8 == 0): ragged last band, non-square planes, 6 bands of 15 rows, x64, one row
    ((1, 2, 13, 9), dict(scale_factor=16)), ((2, 2, 10, 14), dict(scale_factor=8)), ((1, 1, 88, 88), dict(scale_factor=8)),
    ((1, 1, 3, 5), dict(scale_factor=64)), ((2, 1, 1, 4), dict(scale_factor=8)), ((1, 3, 22, 22), dict(scale_factor=32)),
])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_bilinear_fwd_bwd(shape, kw, dtype):
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(1))
    if dtype == torch.bfloat16:
        x = x.bfloat16().float()
    xr = x.clone().requires_grad_(True)
    ref = F.interpolate(xr, mode="bilinear", **kw)
    gout = torch.randn(ref.shape, generator=torch.Generator().manual_seed(2))
    if dtype == torch.bfloat16:
        gout = gout.bfloat16().float()
    ref.backward(gout)
    xd = x.to(DEV, dtype).requires_grad_(True)
    out = P.interpolate_bilinear(xd, **kw)
    assert out.shape == ref.shape and out.dtype == dtype
    out.backward(gout.to(DEV, dtype))
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert (out.float().cpu() - ref).abs().max() <= tol * max(1.0, ref.abs().max().item())
    assert (xd.grad.float().cpu() - xr.grad).abs().max() <= tol * max(1.0, xr.grad.abs().max().item())
    if dtype == torch.float32:   # the explicit float64 numpy restatement agrees too
        sf = kw.get("scale_factor")
        np.testing.assert_allclose(out.detach().cpu().numpy(), O.bilinear_np(x.numpy(), ref.shape[2], ref.shape[3], kw.get("align_corners", False), sf),
                                   rtol=1e-4, atol=5e-5)


@pytest.mark.parametrize("B,C,h,scale,softmax", [(2, 1, 11, 0.25, True), (2, 3, 22, 2, True), (2, 9, 14, 2, True), (1, 4, 16, 2, False),
                                                   (2, 1, 44, 2, False), (1, 9, 56, 2, True)])
def test_dsra_fuse(B, C, h, scale, softmax):
    g = torch.Generator().manual_seed(3)
    dh = int(round(h / scale))
    fg = torch.randn(B, C, h, h, generator=g)
    dfg, dbg = torch.randn(B, C, dh, dh, generator=g), torch.randn(B, C, dh, dh, generator=g)
    gout = torch.randn(B, C, h, h, generator=g)
    r = [t.clone().requires_grad_(True) for t in (fg, dfg, dbg)]
    ref = O.dsra_fuse(r[0], O.interp(r[1], scale), O.interp(r[2], scale), softmax)
    ref.backward(gout)
    d = [t.to(DEV).requires_grad_(True) for t in (fg, dfg, dbg)]
    out = P.dsra_fuse(d[0], d[1], d[2], softmax, scale_factor=scale)
    out.backward(gout.to(DEV))
    assert (out.cpu() - ref).abs().max() <= 1e-5 * max(1.0, ref.abs().max().item())
    for a, b in zip(d, r):
        assert (a.grad.cpu() - b.grad).abs().max() <= 1e-5 * max(1.0, b.grad.abs().max().item())
    if C == 1 and softmax:   # softmax over one channel == 1 -> out = 2*fg, ZERO grad to the deeper maps
        assert torch.equal(out, 2 * d[0].detach())
        assert d[1].grad.abs().max().item() == 0.0 and d[2].grad.abs().max().item() == 0.0


@pytest.mark.parametrize("B,C,h", [(2, 2048, 11), (2, 1024, 22), (1, 512, 44), (1, 7, 5), (3, 12, 11), (2, 16, 7), (16, 2048, 11)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_ra_v1_scale(B, C, h, dtype):
    g = torch.Generator().manual_seed(4)
    x = torch.relu(torch.randn(B, C, h, h, generator=g))
    crop = 2 * torch.randn(B, 1, h, h, generator=g)
    gout = torch.randn(B, C, h, h, generator=g)
    if dtype == torch.bfloat16:
        x, gout = x.bfloat16().float(), gout.bfloat16().float()
    xr, cr = x.clone().requires_grad_(True), crop.clone().requires_grad_(True)
    ref = O.ra_v1_scale(cr, xr)
    ref.backward(gout)
    xd, cd = x.to(DEV, dtype).requires_grad_(True), crop.to(DEV).requires_grad_(True)
    out = P.ra_v1_scale(xd, cd)
    out.backward(gout.to(DEV, dtype))
    tol = 2e-5 if dtype == torch.float32 else 1e-2
    assert (out.float().cpu() - ref).abs().max() <= tol * max(1.0, ref.abs().max().item())
    assert (xd.grad.float().cpu() - xr.grad).abs().max() <= tol * max(1.0, xr.grad.abs().max().item())
    assert (cd.grad.cpu() - cr.grad).abs().max() <= tol * max(1.0, cr.grad.abs().max().item())


@pytest.mark.parametrize("name", list(G.MC_LOSS_CASES))
def test_mc_dual_loss_golden(name):
    case = G.MC_LOSS_CASES[name]
    g = G.load(name)
    P_fg, P_bg, labels = G.mc_loss_inputs(name)
    fg = [t.to(DEV).requires_grad_(True) for t in P_fg]
    bg = [t.to(DEV).requires_grad_(True) for t in P_bg]
    loss = P.mc_dual_loss(fg, bg, labels.to(DEV), case["num_class"])
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= 1e-4 * abs(float(g["loss"]))
    for i in range(4):
        for got, ref in ((fg[i].grad, g[f"dfg{i}"]), (bg[i].grad, g[f"dbg{i}"])):
            assert np.abs(got.cpu().numpy() - ref).max() <= 1e-3 * np.abs(ref).max()


@pytest.mark.parametrize("B,C,H,W,supervision,nmaps", [(4, 9, 224, 224, "mutation", 4), (2, 4, 100, 75, "mutation", 4),
                                                        (2, 9, 64, 64, "deep_supervision", 4), (2, 5, 40, 40, "mutation", 3)])
def test_mc_dual_loss_vs_oracle(B, C, H, W, supervision, nmaps):
    fg = [synth.logits((B, C, H, W), 3, f"f{i}", 1.5) for i in range(nmaps)]
    bg = [synth.logits((B, C, H, W), 3, f"b{i}", 1.5) for i in range(nmaps)]
    labels = synth.class_labels(B, H, W, C, 3)
    rf = [t.clone().requires_grad_(True) for t in fg]
    rb = [t.clone().requires_grad_(True) for t in bg]
    subsets = O.powerset_subsets(nmaps) if supervision == "mutation" else [[i] for i in range(nmaps)]
    ref = O.mc_dual_loss(rf, rb, labels, C, subsets)
    ref.backward()
    df = [t.to(DEV).requires_grad_(True) for t in fg]
    db = [t.to(DEV).requires_grad_(True) for t in bg]
    loss = P.mc_dual_loss(df, db, labels.to(DEV), C, supervision=supervision)
    loss.backward()
    assert abs(loss.item() - ref.item()) <= 1e-4 * abs(ref.item())
    for a, b in zip(df + db, rf + rb):
        assert (a.grad.cpu() - b.grad).abs().max() <= 1e-3 * b.grad.abs().max()


def test_bilinear_multi_matches_single():
    """pv2_bilinear_multi_fwd/bwd (8 maps, one launch) == 8 single-map launches == F.interpolate."""
    from pranet_v2_b200 import engine as E
    B, S = 2, 96
    g = torch.Generator().manual_seed(11)
    lows = [torch.randn(B, 1, S // s, S // s, generator=g) for s in (8, 16, 32, 8, 8, 16, 32, 8)]
    scales = [8, 16, 32, 8, 8, 16, 32, 8]
    gouts = [torch.randn(B, 1, S, S, generator=g) for _ in lows]
    refs, rgrads = [], []
    for x, s, go in zip(lows, scales, gouts):
        xr = x.clone().requires_grad_(True)
        r = F.interpolate(xr, scale_factor=s, mode="bilinear")
        r.backward(go)
        refs.append(r.detach())
        rgrads.append(xr.grad)
    eng = E.Engine(torch.device(DEV), "fp32", True, True)
    maps = [E.Map(x.to(DEV)) for x in lows]
    outs = eng.resize_multi(maps, scales)
    for o, go in zip(outs, gouts):
        o.grads.append(go.to(DEV))
    eng.backward()
    for o, r, m, rg in zip(outs, refs, maps, rgrads):
        assert (o.t.cpu() - r).abs().max() <= 1e-5 * max(1.0, r.abs().max().item())
        assert (m.grad().cpu() - rg).abs().max() <= 1e-5 * max(1.0, rg.abs().max().item())


@pytest.mark.parametrize("shape", [(16, 1, 352, 352), (3, 1, 96, 80), (2, 2, 40, 56), (2, 1, 37, 53)])
def test_structure_loss_fused_equals_two_pass(shape):
    """The two forwards of the loss -- (a) fused: summed-area-table boundary weight inside the loss kernel; (b) two-pass:
    pv2_structure_loss_prepare (boundary_weight_kernel, what a training step runs ahead of the backbone) followed by the
    streaming pv2_structure_loss_fwd_prepared -- write the same 16-bit weight map semantics: losses agree to 2e-6 relative,
    gradients (the backward reads whichever forward's workspace) to 2e-6 of their max; both hold the oracle tolerances."""
    B, Cc, H, W = shape
    g = torch.Generator().manual_seed(3)
    pred, pbg = torch.randn(shape, generator=g) * 3, torch.randn(shape, generator=g) * 3
    mask = synth.ellipse_masks(B * Cc, H, W, 7).reshape(shape)
    if W % 8:
        mask = F.interpolate(mask, size=(H, W), mode="bilinear", align_corners=True)
    res = []
    for two_pass in (False, True):
        p = pred.to(DEV).requires_grad_(True)
        q = pbg.to(DEV).requires_grad_(True)
        md = mask.to(DEV)
        n0 = P._lib.launch_count()
        prepared = P.ops.structure_loss_prepare(md) if two_pass else None
        loss = P.structure_loss_multi([(p, q)], md, prepared=prepared)[0]
        launches = P._lib.launch_count() - n0
        assert launches == (2 if (two_pass or W % 4) else 1), launches      # non-vectorisable widths always take the two-kernel forward
        loss.backward()
        res.append((loss.item(), p.grad.cpu(), q.grad.cpu()))
    ref_p = pred.clone().requires_grad_(True)
    ref_q = pbg.clone().requires_grad_(True)
    rl = O.structure_loss(ref_p, ref_q, mask, 1 - mask)
    rl.backward()
    for l, gp, gq in res:
        assert abs(l - rl.item()) <= 1e-4 * abs(rl.item())
        assert (gp - ref_p.grad).abs().max() <= 1e-3 * ref_p.grad.abs().max()
        assert (gq - ref_q.grad).abs().max() <= 1e-3 * ref_q.grad.abs().max()
    assert abs(res[0][0] - res[1][0]) <= 2e-6 * abs(res[0][0])
    assert (res[0][1] - res[1][1]).abs().max() <= 2e-6 * res[0][1].abs().max()
    assert (res[0][2] - res[1][2]).abs().max() <= 2e-6 * res[0][2].abs().max()


def test_structure_loss_checkerboard_weight_quantisation():
    """Worst case for the 16-bit fixed-point boundary weight (structure_loss.cu: quantisation step 7.6e-5 of a weight in [1, 6]):
    an all-boundary checkerboard mask, where |avgpool31(m) - m| ~ 0.5 on EVERY pixel, and a soft (non-binary) mask on top of it.
    The loss must still hold 1e-4 relative and the gradients 1e-3 of their max against the fp32 reference formula."""
    B, H, W = 4, 352, 352
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    checker = ((yy + xx) % 2).float()[None, None].repeat(B, 1, 1, 1)
    soft = 0.5 * checker + 0.5 * torch.rand(B, 1, H, W, generator=torch.Generator().manual_seed(9))
    g = torch.Generator().manual_seed(4)
    pred, pbg = torch.randn(B, 1, H, W, generator=g) * 3, torch.randn(B, 1, H, W, generator=g) * 3
    for mask in (checker, soft):
        for two_pass in (False, True):
            p = pred.to(DEV).requires_grad_(True)
            q = pbg.to(DEV).requires_grad_(True)
            md = mask.to(DEV)
            loss = P.structure_loss_multi([(p, q)], md, prepared=P.ops.structure_loss_prepare(md) if two_pass else None)[0]
            loss.backward()
            rp, rq = pred.clone().requires_grad_(True), pbg.clone().requires_grad_(True)
            rl = O.structure_loss(rp, rq, mask, 1 - mask)
            rl.backward()
            assert abs(loss.item() - rl.item()) <= 1e-4 * abs(rl.item()), (loss.item(), rl.item())
            assert (p.grad.cpu() - rp.grad).abs().max() <= 1e-3 * rp.grad.abs().max()
            assert (q.grad.cpu() - rq.grad).abs().max() <= 1e-3 * rq.grad.abs().max()


# ------------------------------------------------------------------------------------------------
# loss from the low-resolution maps
# ------------------------------------------------------------------------------------------------
def _lowres_case(B, Cc, H, W, scales, seed, soft=False):
    maps = []
    for k, s in enumerate(scales):
        shape = (B, Cc, H // s, W // s)
        maps.append((synth.logits(shape, seed, f"lf{k}"), synth.logits(shape, seed, f"lb{k}")))
    m = (synth.soft_masks(B * Cc, H, W, seed) if soft else synth.ellipse_masks(B * Cc, H, W, seed)).view(B, Cc, H, W)
    return maps, m


def _oracle_lowres(maps, scales, m, mb, wts):
    """pranet.py:349-415 final upsamples + MyTrain_med.py:78-82 losses on CPU: F.interpolate -> oracle structure_loss."""
    leaves = [(a.clone().requires_grad_(True), b.clone().requires_grad_(True)) for a, b in maps]
    losses = torch.stack([O.structure_loss(O.interp(a, scale=s), O.interp(b, scale=s), m, mb) for (a, b), s in zip(leaves, scales)])
    (losses * wts).sum().backward()
    return losses.detach(), leaves


@pytest.mark.parametrize("B,Cc,H,W,scales,soft,pass_bg", [
    (2, 1, 352, 352, (8, 16, 32, 8), False, False),       # PraNet-V2 geometry (pranet.py:349-350,370-371,392-393,414-415)
    (2, 1, 256, 256, (8, 16, 32, 8), True, False),        # multi-scale rate 0.75: soft masks (MyTrain_med.py:70-73)
    (1, 3, 224, 224, (4, 8, 16, 32), False, True),        # EMCAD scales (networks.py:116-123), explicit mask_bg
    (2, 2, 96, 132, (4, 12), True, True),                 # ragged: W % 128 != 0, partial tiles, 2 scales
    (1, 1, 64, 64, (16,), False, False),                  # one scale, one tile row pair
])
def test_structure_loss_lowres_vs_oracle(B, Cc, H, W, scales, soft, pass_bg):
    maps, m = _lowres_case(B, Cc, H, W, scales, 21, soft)
    mb = 1 - m
    wts = torch.tensor([1.0, 0.5, 2.0, 1.5][:len(scales)])
    ref_losses, leaves = _oracle_lowres(maps, scales, m, mb, wts)
    dev = [(a.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)) for a, b in maps]
    n0 = P._lib.launch_count()
    losses = P.structure_loss_lowres(dev, list(scales), m.to(DEV), mb.to(DEV) if pass_bg else None)
    (losses * wts.to(DEV)).sum().backward()
    assert P._lib.launch_count() - n0 == 3                 # fused forward, fused backward, fold: no bilinear / full-res launches
    assert (losses.cpu() - ref_losses).abs().max().item() <= 1e-4 * ref_losses.abs().max().item()
    for (a, b), (ra, rb) in zip(dev, leaves):
        for got, want in ((a.grad, ra.grad), (b.grad, rb.grad)):
            assert torch.isfinite(got).all()
            assert (got.cpu() - want).abs().max().item() <= 1e-3 * want.abs().max().item() + 1e-12


@pytest.mark.parametrize("name", list(G.LOWRES_LOSS_CASES))
def test_structure_loss_lowres_golden(name):
    """Fused kernels == the UNMODIFIED reference's statements (final F.interpolate calls of pranet.py + MyTrain_med.py:74,78-82,
    frozen by oracle/make_golden.py gen_lowres_loss): the four losses, their sum, and the gradients of the eight low-res maps."""
    g = G.load(name)
    maps, m = G.lowres_loss_inputs(name)
    dev = [t.to(DEV).requires_grad_(True) for t in maps]
    n0 = P._lib.launch_count()
    losses = P.structure_loss_lowres([(dev[i], dev[i + 4]) for i in range(4)], G.lowres_loss_scales(name), m.to(DEV))
    losses.sum().backward()
    assert P._lib.launch_count() - n0 == 3                  # every golden geometry is inside the fused kernels' coverage
    np.testing.assert_allclose(losses.detach().cpu().numpy(), g["losses"], rtol=1e-4)
    assert abs(losses.sum().item() - float(g["loss"])) <= 1e-4 * abs(float(g["loss"]))
    for i, t in enumerate(dev):
        ref = g[f"d{i}"]
        np.testing.assert_allclose(t.grad.cpu().numpy(), ref, rtol=0, atol=1e-3 * np.abs(ref).max())


def test_structure_loss_lowres_equals_unfused_full_size():
    """B = 16 x 352^2, four scales: the fused path against the module path's own kernels (bilinear x8 -> structure_loss x4 ->
    bilinear backward) -- same taps, same weight map, so agreement is at fp32 summation-order level; and it is deterministic."""
    scales = (8, 16, 32, 8)
    maps, m = _lowres_case(16, 1, 352, 352, scales, 4)
    md = m.to(DEV)
    wts = torch.tensor([1.0, 1.0, 1.0, 1.0], device=DEV)
    res = []
    for fused in (True, False, True):
        dev = [(a.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)) for a, b in maps]
        if fused:
            losses = P.structure_loss_lowres(dev, list(scales), md)
        else:
            up = [(P.interpolate_bilinear(a, scale_factor=s), P.interpolate_bilinear(b, scale_factor=s)) for (a, b), s in zip(dev, scales)]
            losses = P.structure_loss_multi(up, md)
        (losses * wts).sum().backward()
        res.append((losses.detach().cpu(), [t.grad.cpu() for pair in dev for t in pair]))
    (lf, gf), (lu, gu), (lf2, gf2) = res
    assert (lf - lu).abs().max().item() <= 2e-6 * lu.abs().max().item()
    for a, b in zip(gf, gu):
        assert (a - b).abs().max().item() <= 2e-5 * b.abs().max().item()
    assert torch.equal(lf, lf2) and all(torch.equal(a, b) for a, b in zip(gf, gf2))      # no atomics: bit-reproducible


def test_structure_loss_lowres_unsupported_geometry_takes_unfused_kernels():
    """x2 upsampling is outside the fused kernels' coverage: the op routes through interpolate_bilinear + structure_loss_multi
    (still pv2 kernels, more launches) and matches the oracle; CPU tensors raise like every other op."""
    maps, m = _lowres_case(2, 1, 64, 64, (2,), 8)
    ref_losses, leaves = _oracle_lowres(maps, (2,), m, 1 - m, torch.ones(1))
    dev = [(a.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)) for a, b in maps]
    n0 = P._lib.launch_count()
    losses = P.structure_loss_lowres(dev, [2], m.to(DEV))
    losses.sum().backward()
    assert P._lib.launch_count() - n0 > 3
    assert (losses.cpu() - ref_losses).abs().max().item() <= 1e-4 * ref_losses.abs().max().item()
    assert (dev[0][0].grad.cpu() - leaves[0][0].grad).abs().max().item() <= 1e-3 * leaves[0][0].grad.abs().max().item()
    with pytest.raises(RuntimeError, match="CUDA-only"):
        P.structure_loss_lowres(maps, [2], m)
    with pytest.raises(ValueError):
        P.structure_loss_lowres(dev, [4], m.to(DEV))       # 32 * 4 != 64
    # the C ABI itself refuses what it does not cover, loudly
    lib = P._lib.load()
    a, b = dev[0][0].detach(), dev[0][1].detach()
    md = m.to(DEV)
    ws_bytes = lib.pv2_structure_loss_lowres_workspace_bytes(2, 64, 64, 1)
    ws = torch.empty(ws_bytes // 4, device=DEV)
    loss = torch.empty(1, device=DEV)
    pf, k1 = P._lib.ptr_array([a]); pb, k2 = P._lib.ptr_array([b])
    ph, k3 = P._lib.int_array([32]); pw, k4 = P._lib.int_array([32])
    pr, k5 = P._lib.float_array([0.5])
    st = lib.pv2_structure_loss_lowres_fwd(pf, pb, ph, pw, pr, pr, md.data_ptr(), None, 1, 2, 64, 64, loss.data_ptr(), ws.data_ptr(), ws_bytes,
                                           torch.cuda.current_stream().cuda_stream)
    assert st != 0 and b"up-scaling by >= 4" in lib.pv2_last_error()


@pytest.mark.parametrize("N,H,W,Cc,nslabs", [(2, 11, 11, 32, 1), (16, 22, 22, 64, 3), (2, 22, 18, 96, 2), (1, 2, 3, 4, 1), (3, 9, 17, 20, 2), (1, 44, 44, 32, 1)])
def test_up2_nhwc_bwd_vs_aten(N, H, W, Cc, nslabs):
    """pv2_up2_nhwc_bwd (backward of nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True), pranet.py:93, on NHWC rows with
    the upstream gradient arriving as several slabs that are summed on load) against ATen's backward: the shared-memory tiled kernel
    (ragged tiles, partial channel groups, slabs wider than C with a channel offset)."""
    import ctypes
    lib = P._lib.load()
    g = torch.Generator().manual_seed(H * 100 + W)
    slabs, ref_g = [], torch.zeros(N, Cc, 2 * H, 2 * W)
    for i in range(nslabs):
        ld, off = Cc + 8 * i, 4 * i                      # slab i holds the gradient in channels [off, off + C) of rows ld wide
        t = torch.randn(N * 2 * H * 2 * W, ld, generator=g)
        slabs.append((t.to(DEV), ld, off))
        ref_g += t[:, off:off + Cc].reshape(N, 2 * H, 2 * W, Cc).permute(0, 3, 1, 2)
    x = torch.zeros(N, Cc, H, W, requires_grad=True)
    F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True).backward(ref_g)
    din = torch.full((N * H * W, Cc), float("nan"), device=DEV)
    pp, keep = P._lib.ptr_array([t for t, _, _ in slabs])
    lds = (ctypes.c_int * nslabs)(*[ld for _, ld, _ in slabs])
    offs = (ctypes.c_int * nslabs)(*[off for _, _, off in slabs])
    P._lib.check(lib.pv2_up2_nhwc_bwd(pp, lds, offs, nslabs, din.data_ptr(), Cc, N, H, W, Cc, torch.cuda.current_stream().cuda_stream), "pv2_up2_nhwc_bwd")
    got = din.cpu().reshape(N, H, W, Cc).permute(0, 3, 1, 2)
    assert torch.isfinite(got).all()
    assert (got - x.grad).abs().max().item() <= 2e-5 * max(1.0, x.grad.abs().max().item())
