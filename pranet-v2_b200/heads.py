"""The DSRA decoder head as drop-in nn.Modules.

Parameter containers (and therefore state_dict keys) are exactly the reference's -- `X.conv.weight`,
`X.bn.{weight,bias,running_mean,running_var,num_batches_tracked}` for every BasicConv2d, `branch{b}.{i}`
inside RFB_modified, `conv_upsample{1..5} / conv_concat{2,3} / conv4 / conv5_fg / conv5_bg` inside
aggregation (binary_seg/lib/pranet.py:31-125) -- so `RES-V2.pth` / `PVT-V2.pth` load unchanged.  The modules own
no computation of their own: `forward` hands a *runner* to `engine.run_head`, and the `*_run` functions below
describe each block in terms of engine ops (tcgen05 convs, BN statistics / apply kernels, NHWC upsample ...).
Every block is also callable on its own with NCHW tensors, like the reference modules.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import engine as E


def _params(*mods):
    out = []
    for m in mods:
        out += [p for p in m.parameters()]
    return out


def _sink(in_grads, i):
    def f(g):
        in_grads[i] = g if in_grads[i] is None else in_grads[i] + g
    return f


class BasicConv2d(nn.Module):
    """conv (bias-free) -> BatchNorm, no ReLU (binary_seg/lib/pranet.py:31-43; the `relu` member of the
    reference is never applied in its forward and has no parameters, so it is not reproduced)."""

    def __init__(self, in_planes, out_planes, kernel_size, stride=1, padding=0, dilation=1):
        super().__init__()
        self.conv = nn.Conv2d(in_planes, out_planes, kernel_size=kernel_size, stride=stride,
                              padding=padding, dilation=dilation, bias=False)
        self.bn = nn.BatchNorm2d(out_planes)

    def run(self, eng, x: E.Act, relu=False, out=None, out_map=False):
        raw = eng.conv(x, [self.conv], [self.bn])
        return eng.bn_apply((raw, 0, self.conv.out_channels, self.bn, None), relu=relu, out=out, out_map=out_map)

    def forward(self, x, relu: bool = False):
        def runner(eng, inputs, in_grads):
            return [self.run(eng, eng.from_nchw(inputs[0], _sink(in_grads, 0)), relu=relu, out_map=True)]
        return E.run_head(runner, [x], _params(self), self.training)[0]


def rfb_begin(eng, rfb: "RFB_modified", raw1x1: E.Raw):
    """The concat buffer the four RFB branches write their slices of."""
    c = rfb.conv_res.conv.out_channels
    return eng.concat_buffer(raw1x1.N, raw1x1.H, raw1x1.W, [c] * 4)


def rfb_branch(eng, rfb: "RFB_modified", state, raw1x1: E.Raw, off: int, b: int):
    """Branch b (0..3) of RFB_modified.forward (pranet.py:76-79): independent of the other branches."""
    c = rfb.conv_res.conv.out_channels
    _, sl = state
    if b == 0:
        eng.bn_apply((raw1x1, off, c, rfb.branch0[0].bn, None), out=sl[0])
        return
    br = getattr(rfb, f"branch{b}")
    t = eng.bn_apply((raw1x1, off + b * c, c, br[0].bn, None))
    t = br[1].run(eng, t)
    t = br[2].run(eng, t)
    br[3].run(eng, t, out=sl[b])


def rfb_finish(eng, rfb: "RFB_modified", state, raw1x1: E.Raw, off: int, out_map=False):
    """conv_cat over the concat, + conv_res, ReLU (pranet.py:80-82)."""
    c = rfb.conv_res.conv.out_channels
    whole, _ = state
    raw_cat = eng.conv(whole, [rfb.conv_cat.conv], [rfb.conv_cat.bn])
    return eng.bn_apply((raw_cat, 0, c, rfb.conv_cat.bn, None), src2=(raw1x1, off + 4 * c, c, rfb.conv_res.bn, None), combine=1, relu=True,
                        out_map=out_map)


def rfb_run(eng, rfb: "RFB_modified", raw1x1: E.Raw, off: int, out_map=False):
    """RFB_modified.forward (pranet.py:75-83) given the raw output of the horizontally fused 1x1 convs
    [branch0.0 | branch1.0 | branch2.0 | branch3.0 | conv_res] starting at channel `off` of `raw1x1`.  The four branches
    run as a parallel section when the engine has side streams."""
    state = rfb_begin(eng, rfb, raw1x1)
    eng.fork(4)
    for b in range(4):
        with eng.branch(b + 1):
            rfb_branch(eng, rfb, state, raw1x1, off, b)
    eng.join()
    return rfb_finish(eng, rfb, state, raw1x1, off, out_map)


def rfb_convs(rfb: "RFB_modified"):
    return [rfb.branch0[0].conv, rfb.branch1[0].conv, rfb.branch2[0].conv, rfb.branch3[0].conv, rfb.conv_res.conv]


def rfb_bns(rfb: "RFB_modified"):
    return [rfb.branch0[0].bn, rfb.branch1[0].bn, rfb.branch2[0].bn, rfb.branch3[0].bn, rfb.conv_res.bn]


class RFB_modified(nn.Module):
    """Receptive-field block (pranet.py:46-83): five 1x1 reductions of the same input (ONE fused GEMM here), three
    (1xk, kx1, 3x3 dil k) chains, 3x3 over the concat, residual add, ReLU."""

    def __init__(self, in_channel, out_channel):
        super().__init__()
        self.branch0 = nn.Sequential(BasicConv2d(in_channel, out_channel, 1))
        for b, k in ((1, 3), (2, 5), (3, 7)):
            setattr(self, f"branch{b}", nn.Sequential(
                BasicConv2d(in_channel, out_channel, 1),
                BasicConv2d(out_channel, out_channel, kernel_size=(1, k), padding=(0, k // 2)),
                BasicConv2d(out_channel, out_channel, kernel_size=(k, 1), padding=(k // 2, 0)),
                BasicConv2d(out_channel, out_channel, 3, padding=k, dilation=k)))
        self.conv_cat = BasicConv2d(4 * out_channel, out_channel, 3, padding=1)
        self.conv_res = BasicConv2d(in_channel, out_channel, 1)

    def forward(self, x):
        def runner(eng, inputs, in_grads):
            a = eng.from_nchw(inputs[0], _sink(in_grads, 0))
            return [rfb_run(eng, self, eng.conv(a, rfb_convs(self), rfb_bns(self)), 0, out_map=True)]
        return E.run_head(runner, [x], _params(self), self.training)[0]


def aggregation_run(eng, agg: "aggregation", x1: E.Act, x2: E.Act, x3: E.Act):
    """aggregation.forward (pranet.py:109-125): x1 deepest (H/32), x2 (H/16), x3 (H/8).  Returns the list of
    head Maps ([fg, bg] for V2, [single] for V1)."""
    c = x1.C
    cat2, s2 = eng.concat_buffer(x2.N, x2.H, x2.W, [c, c])
    cat3, s3 = eng.concat_buffer(x3.N, x3.H, x3.W, [c, 2 * c])
    # three independent chains (pranet.py:109-118): [up(x1) -> conv_upsample1|4 -> x2_1, cat2 -> conv_concat2 -> up -> conv_upsample5],
    # [up(up(x1)) -> conv_upsample2] and [up(x2) -> conv_upsample3]; they meet in x3_1 = conv_upsample2(.) * conv_upsample3(.) * x3
    res = {}
    eng.fork(3)
    with eng.branch(1):
        up_x1 = eng.up2(x1)
        # conv_upsample1 and conv_upsample4 read the same tensor: one GEMM
        raw_a = eng.conv(up_x1, [agg.conv_upsample1.conv, agg.conv_upsample4.conv], [agg.conv_upsample1.bn, agg.conv_upsample4.bn])
        eng.bn_apply((raw_a, 0, c, agg.conv_upsample1.bn, None), mult=x2, out=s2[0])                   # x2_1
        eng.bn_apply((raw_a, c, c, agg.conv_upsample4.bn, None), out=s2[1])
        x2_2 = agg.conv_concat2.run(eng, cat2)
        agg.conv_upsample5.run(eng, eng.up2(x2_2), out=s3[1])
    with eng.branch(2):
        res["b"] = eng.conv(eng.up2(eng.up2(x1)), [agg.conv_upsample2.conv], [agg.conv_upsample2.bn])
    with eng.branch(3):
        res["c"] = eng.conv(eng.up2(x2), [agg.conv_upsample3.conv], [agg.conv_upsample3.bn])
    eng.join()
    eng.bn_apply((res["b"], 0, c, agg.conv_upsample2.bn, None), src2=(res["c"], 0, c, agg.conv_upsample3.bn, None),
                 combine=2, mult=x3, out=s3[0])                                                         # x3_1
    x3_2 = agg.conv_concat3.run(eng, cat3)
    x = agg.conv4.run(eng, x3_2)
    heads = [agg.conv5] if hasattr(agg, "conv5") else [agg.conv5_fg, agg.conv5_bg]
    raw_h = eng.conv(x, heads)                                                                         # biased 1x1 heads, fused
    outs, o = [], 0
    for h in heads:
        outs.append(eng.bn_apply((raw_h, o, h.out_channels, None, h.bias), out_map=True))
        o += h.out_channels
    return outs


class aggregation(nn.Module):
    """Partial decoder (pranet.py:86-125; V1: PraNet_Res2Net.py:64-98).  num_class=None builds the V1
    single-head variant (`conv5`), otherwise the V2 fg/bg pair (`conv5_fg`, `conv5_bg`, 1x1 with bias)."""

    def __init__(self, channel, num_class=None):
        super().__init__()
        c = channel
        self.conv_upsample1 = BasicConv2d(c, c, 3, padding=1)
        self.conv_upsample2 = BasicConv2d(c, c, 3, padding=1)
        self.conv_upsample3 = BasicConv2d(c, c, 3, padding=1)
        self.conv_upsample4 = BasicConv2d(c, c, 3, padding=1)
        self.conv_upsample5 = BasicConv2d(2 * c, 2 * c, 3, padding=1)
        self.conv_concat2 = BasicConv2d(2 * c, 2 * c, 3, padding=1)
        self.conv_concat3 = BasicConv2d(3 * c, 3 * c, 3, padding=1)
        self.conv4 = BasicConv2d(3 * c, 3 * c, 3, padding=1)
        if num_class is None:
            self.conv5 = nn.Conv2d(3 * c, 1, 1)
        else:
            self.conv5_fg = nn.Conv2d(3 * c, num_class, 1)
            self.conv5_bg = nn.Conv2d(3 * c, num_class, 1)

    def forward(self, x1, x2, x3):
        def runner(eng, inputs, in_grads):
            acts = [eng.from_nchw(t, _sink(in_grads, i)) for i, t in enumerate(inputs)]
            return aggregation_run(eng, self, *acts)
        outs = E.run_head(runner, [x1, x2, x3], _params(self), self.training)
        return outs[0] if len(outs) == 1 else tuple(outs)


# ----------------------------------------------------------------------------------------------------------
# DSRA stages of the multiclass host decoders
# ----------------------------------------------------------------------------------------------------------
def dual_heads_run(eng, fg_mod, bg_mod, feat: E.Act):
    """fg / bg heads on the same decoder feature as ONE GEMM (N = 2*num_class).  BasicConv2d heads (conv + BN:
    EMCAD/lib/decoders.py:434-444, MERIT/lib/decoders.py:298-322) or biased 1x1 nn.Conv2d (MIST/lib/MIST.py:403-412)."""
    if isinstance(fg_mod, BasicConv2d):
        raw = eng.conv(feat, [fg_mod.conv, bg_mod.conv], [fg_mod.bn, bg_mod.bn])
        c = fg_mod.conv.out_channels
        return (eng.bn_apply((raw, 0, c, fg_mod.bn, None), out_map=True), eng.bn_apply((raw, c, c, bg_mod.bn, None), out_map=True))
    raw = eng.conv(feat, [fg_mod, bg_mod])
    c = fg_mod.out_channels
    return (eng.bn_apply((raw, 0, c, None, fg_mod.bias), out_map=True), eng.bn_apply((raw, c, c, None, bg_mod.bias), out_map=True))


class DSRAStages(nn.Module):
    """The DSRA part of EMCAD_dual / CASCADE_Add_dual / CAM as one plug-in: owns `<name>_fg` / `<name>_bg` for
    the four stages (registered on the PARENT under the reference's attribute names by `attach`) and runs
        fg_k, bg_k = heads(d_k);  fg_k <- fg_k + fg_k * softmax_c(resize(fg_{k+1}) - resize(bg_{k+1}))
    deep -> shallow, returning [d4_fg, d3_fg, d2_fg, d1_fg, d4_bg, d3_bg, d2_bg, d1_bg]
    (EMCAD/lib/decoders.py:454-526; MERIT/lib/decoders.py:342-431; MIST/lib/MIST.py:418-451)."""

    def __init__(self, parent: nn.Module, channels, num_class, names=("ConvBlock4", "ConvBlock3", "ConvBlock2", "ConvBlock1"),
                 kernel_sizes=(1, 3, 3, 3), bn=True, use_softmax=True):
        super().__init__()
        self.use_softmax = use_softmax
        mods = []
        for c, n, k in zip(channels, names, kernel_sizes):
            if bn:
                fg, bg = BasicConv2d(c, num_class, k, padding=k // 2), BasicConv2d(c, num_class, k, padding=k // 2)
            else:
                fg, bg = nn.Conv2d(c, num_class, 1), nn.Conv2d(c, num_class, 1)
            setattr(parent, n + "_fg", fg)
            setattr(parent, n + "_bg", bg)
            mods.append((fg, bg))
        object.__setattr__(self, "_mods", mods)      # owned (registered) by the parent, not by this helper

    def forward(self, feats):
        mods = self._mods

        def runner(eng, inputs, in_grads):
            fgs, bgs = [], []
            for i, (x, (fg_m, bg_m)) in enumerate(zip(inputs, mods)):
                fg, bg = dual_heads_run(eng, fg_m, bg_m, eng.from_nchw(x, _sink(in_grads, i)))
                if fgs:
                    fg = eng.fuse(fg, fgs[-1], bgs[-1], self.use_softmax, None)
                fgs.append(fg)
                bgs.append(bg)
            return fgs + bgs
        params = []
        for fg_m, bg_m in mods:
            params += _params(fg_m, bg_m)
        return list(E.run_head(runner, list(feats), params, mods[0][0].training))
