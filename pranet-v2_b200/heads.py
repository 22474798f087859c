"""The DSRA decoder head as drop-in nn.Modules.

Parameter containers (and therefore state_dict keys) are exactly the reference's -- `X.conv.weight`,
`X.bn.{weight,bias,running_mean,running_var,num_batches_tracked}` for every BasicConv2d, `branch{b}.{i}`
inside RFB_modified, `conv_upsample{1..5} / conv_concat{2,3} / conv4 / conv5_fg / conv5_bg` inside
aggregation (binary_seg/lib/pranet.py:31-125) -- so `RES-V2.pth` / `PVT-V2.pth` load unchanged.  What the
forward passes *do* is dispatched to the pv2 kernels through `engine` / `ops`.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import engine, ops


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


class BasicConv2d(nn.Module):
    """conv (bias-free) -> BatchNorm, no ReLU (binary_seg/lib/pranet.py:31-43; the `relu` member of the
    reference is never applied in its forward and has no parameters, so it is not reproduced)."""

    def __init__(self, in_planes, out_planes, kernel_size, stride=1, padding=0, dilation=1):
        super().__init__()
        self.conv = nn.Conv2d(in_planes, out_planes, kernel_size=kernel_size, stride=stride,
                              padding=padding, dilation=dilation, bias=False)
        self.bn = nn.BatchNorm2d(out_planes)

    def forward(self, x, relu: bool = False):
        return engine.conv_bn_act(x, self.conv, self.bn, relu)


class RFB_modified(nn.Module):
    """Receptive-field block (pranet.py:46-83): five 1x1 reductions of the same input, three
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
        outs = [self.branch0(x), self.branch1(x), self.branch2(x), self.branch3(x)]
        return engine.add_relu(self.conv_cat(engine.concat(outs)), self.conv_res(x))


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

    def trunk(self, x1, x2, x3):
        up = engine.up2_align_corners
        x2_1 = engine.mul(self.conv_upsample1(up(x1)), x2)
        x3_1 = engine.mul(engine.mul(self.conv_upsample2(up(up(x1))), self.conv_upsample3(up(x2))), x3)
        x2_2 = self.conv_concat2(engine.concat([x2_1, self.conv_upsample4(up(x1))]))
        x3_2 = self.conv_concat3(engine.concat([x3_1, self.conv_upsample5(up(x2_2))]))
        return self.conv4(x3_2)

    def forward(self, x1, x2, x3):
        x = self.trunk(x1, x2, x3)
        if hasattr(self, "conv5"):
            return engine.conv_bias(x, self.conv5)
        return engine.conv_bias(x, self.conv5_fg), engine.conv_bias(x, self.conv5_bg)


class DualHeadStage(nn.Module):
    """One DSRA stage of a multiclass host decoder: fg / bg heads on the same decoder feature, then
    fg <- fg + fg * softmax_c(resize(deeper_fg) - resize(deeper_bg)).

    Mirrors `ConvBlock{k}_fg/_bg` + the fusion lines of EMCAD_dual (EMCAD/lib/decoders.py:434-444,
    454-523) and CASCADE_Add_dual (MERIT/lib/decoders.py:298-322, 342-428); with bn=False the heads are
    the biased 1x1 convs of MIST's CAM (MIST/lib/MIST.py:403-449).  The two head modules are registered
    on the *parent* under the reference's attribute names by `attach_dual_heads`."""

    def __init__(self, fg: nn.Module, bg: nn.Module, use_softmax=True):
        super().__init__()
        object.__setattr__(self, "_fg", fg)   # not registered here: owned by the parent
        object.__setattr__(self, "_bg", bg)
        self.use_softmax = use_softmax

    def forward(self, feat, deeper_fg=None, deeper_bg=None):
        if isinstance(self._fg, BasicConv2d):
            fg, bg = self._fg(feat), self._bg(feat)
        else:
            fg, bg = engine.conv_bias(feat, self._fg), engine.conv_bias(feat, self._bg)
        if deeper_fg is not None:
            fg = ops.dsra_fuse(fg, deeper_fg, deeper_bg, self.use_softmax)
        return fg, bg


def attach_dual_heads(parent: nn.Module, channels, num_class, names=("ConvBlock4", "ConvBlock3", "ConvBlock2", "ConvBlock1"),
                      kernel_sizes=(1, 3, 3, 3), bn=True, use_softmax=True):
    """Registers `<name>_fg` / `<name>_bg` on `parent` (same keys as the reference decoders) and returns
    the list of DualHeadStage callables, deep -> shallow."""
    stages = []
    for c, n, k in zip(channels, names, kernel_sizes):
        if bn:
            fg, bg = BasicConv2d(c, num_class, k, padding=k // 2), BasicConv2d(c, num_class, k, padding=k // 2)
        else:
            fg, bg = nn.Conv2d(c, num_class, 1), nn.Conv2d(c, num_class, 1)
        setattr(parent, n + "_fg", fg)
        setattr(parent, n + "_bg", bg)
        stages.append(DualHeadStage(fg, bg, use_softmax))
    return stages


def dsra_cascade(stages, feats):
    """Runs the DSRA stages over decoder features d4..d1 (deep -> shallow) and returns
    [d4_fg, d3_fg, d2_fg, d1_fg, d4_bg, d3_bg, d2_bg, d1_bg] like EMCAD_dual.forward (decoders.py:526)."""
    fgs, bgs = [], []
    for st, d in zip(stages, feats):
        fg, bg = st(d, fgs[-1] if fgs else None, bgs[-1] if bgs else None)
        fgs.append(fg)
        bgs.append(bg)
    return fgs + bgs
