"""Drop-in host networks: same class names, constructor kwargs, forward signatures, output tuples and
state_dict keys as the reference (binary_seg/lib/pranet.py:129,268; binary_seg/lib/PraNet_Res2Net.py:101,188).

Backbones are stock PyTorch (`backbones.py`, out of scope); everything after the backbone is the DSRA
head running on the pv2 kernels.  `forward_head(x2, x3, x4)` exposes the head alone (used by the
microbenchmarks and the parity tests; the reference has no separable head).

Differences from the reference that are deliberate:
  * no checkpoint is read at construction time (the reference hard-codes `./models/pvt_v2_b2.pth`,
    pranet.py:147-152, and `../models/res2net50_v1b_26w_4s-3cf99910.pth`, Res2Net_v1b.py:198); use
    `load_state_dict` / `load_backbone`.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import engine, ops
from .backbones import PvtV2B2, Res2Net50
from .heads import BasicConv2d, RFB_modified, aggregation

RES2NET_CH = (512, 1024, 2048)
PVT_CH = (128, 320, 512)


class _PraNetBase(nn.Module):
    def _build_head(self, channels, channel, num_class, v1):
        c2, c3, c4 = channels
        self.rfb2_1 = RFB_modified(c2, channel)
        self.rfb3_1 = RFB_modified(c3, channel)
        self.rfb4_1 = RFB_modified(c4, channel)
        self.agg1 = aggregation(channel, None if v1 else num_class)
        # reverse-attention stack on x4: 1x1 then three 5x5
        self.ra4_conv1 = BasicConv2d(c4, 256, kernel_size=1)
        self.ra4_conv2 = BasicConv2d(256, 256, kernel_size=5, padding=2)
        self.ra4_conv3 = BasicConv2d(256, 256, kernel_size=5, padding=2)
        self.ra4_conv4 = BasicConv2d(256, 256, kernel_size=5, padding=2)
        # on x3 / x2: 1x1 then two 3x3
        for s, cin in ((3, c3), (2, c2)):
            setattr(self, f"ra{s}_conv1", BasicConv2d(cin, 64, kernel_size=1))
            setattr(self, f"ra{s}_conv2", BasicConv2d(64, 64, kernel_size=3, padding=1))
            setattr(self, f"ra{s}_conv3", BasicConv2d(64, 64, kernel_size=3, padding=1))
        if v1:
            self.ra4_conv5 = BasicConv2d(256, 1, kernel_size=1)
            self.ra3_conv4 = BasicConv2d(64, 1, kernel_size=3, padding=1)
            self.ra2_conv4 = BasicConv2d(64, 1, kernel_size=3, padding=1)
        else:
            for t in ("fg", "bg"):
                setattr(self, f"ra4_conv5_{t}", BasicConv2d(256, num_class, kernel_size=1))
                setattr(self, f"ra3_conv4_{t}", BasicConv2d(64, num_class, kernel_size=3, padding=1))
                setattr(self, f"ra2_conv4_{t}", BasicConv2d(64, num_class, kernel_size=3, padding=1))

    def _stack(self, stage, x, n):
        """ra{stage}_conv1 without ReLU, then ra{stage}_conv2..n each followed by ReLU
        (pranet.py:357-360, 378-380, 400-402)."""
        x = getattr(self, f"ra{stage}_conv1")(x)
        for i in range(2, n + 1):
            x = getattr(self, f"ra{stage}_conv{i}")(x, relu=True)
        return x

    def load_backbone(self, state_dict):
        """Key-filtered merge, like pranet.py:148-152."""
        bb = self.backbone if hasattr(self, "backbone") else self.resnet
        own = bb.state_dict()
        own.update({k: v for k, v in state_dict.items() if k in own})
        bb.load_state_dict(own)


class _V2Mixin(_PraNetBase):
    def _init_v2(self, channels, channel, num_class, sem_downsample, use_softmax):
        self.idx = range(10)
        self.num_class = num_class
        self.sem_downsample = sem_downsample
        self.use_softmax = use_softmax
        # 1 -> 3 channel stem; PraNet_V2 defines but never uses it, PVT_PraNet_V2 applies it to grayscale input
        self.conv = nn.Sequential(nn.Conv2d(1, 3, kernel_size=1), nn.BatchNorm2d(3), nn.ReLU(inplace=True))
        self._build_head(channels, channel, num_class, v1=False)

    def forward_head(self, x2, x3, x4):
        """pranet.py:343-417 -> (l2_fg, l3_fg, l4_fg, l5_fg, l2_bg, l3_bg, l4_bg, l5_bg)."""
        up = ops.interpolate_bilinear
        sd = self.sem_downsample
        x2_rfb, x3_rfb, x4_rfb = self.rfb2_1(x2), self.rfb3_1(x3), self.rfb4_1(x4)
        ra5_fg, ra5_bg = self.agg1(x4_rfb, x3_rfb, x2_rfb)
        l5_fg, l5_bg = up(ra5_fg, scale_factor=8 / sd), up(ra5_bg, scale_factor=8 / sd)
        # DSRA3: deeper maps are the x0.25-resized coarse maps (resize fused into the fusion kernel)
        t = self._stack(4, x4, 4)
        fg, bg = self.ra4_conv5_fg(t), self.ra4_conv5_bg(t)
        fg = ops.dsra_fuse(fg, ra5_fg, ra5_bg, self.use_softmax, scale_factor=0.25)
        l4_fg, l4_bg = up(fg, scale_factor=32 / sd), up(bg, scale_factor=32 / sd)
        # DSRA2
        t = self._stack(3, x3, 3)
        fg3, bg3 = self.ra3_conv4_fg(t), self.ra3_conv4_bg(t)
        fg3 = ops.dsra_fuse(fg3, fg, bg, self.use_softmax, scale_factor=2)
        l3_fg, l3_bg = up(fg3, scale_factor=16 / sd), up(bg3, scale_factor=16 / sd)
        # DSRA1
        t = self._stack(2, x2, 3)
        fg2, bg2 = self.ra2_conv4_fg(t), self.ra2_conv4_bg(t)
        fg2 = ops.dsra_fuse(fg2, fg3, bg3, self.use_softmax, scale_factor=2)
        l2_fg, l2_bg = up(fg2, scale_factor=8 / sd), up(bg2, scale_factor=8 / sd)
        return l2_fg, l3_fg, l4_fg, l5_fg, l2_bg, l3_bg, l4_bg, l5_bg


class PraNet_V2(_V2Mixin):
    """Res2Net-50 PraNet-V2 (binary_seg/lib/pranet.py:268-417)."""

    def __init__(self, channel=32, num_class=3, sem_downsample=1, use_softmax=True):
        super().__init__()
        self.backbone = Res2Net50()
        self._init_v2(RES2NET_CH, channel, num_class, sem_downsample, use_softmax)

    def forward(self, x, segSize=None):
        _, x2, x3, x4 = self.backbone.pyramid(x)
        return self.forward_head(x2, x3, x4)


class PVT_PraNet_V2(_V2Mixin):
    """PVTv2-b2 PraNet-V2 (pranet.py:129-263)."""

    def __init__(self, channel=32, num_class=3, sem_downsample=1, use_softmax=True):
        super().__init__()
        self.backbone = PvtV2B2()
        self._init_v2(PVT_CH, channel, num_class, sem_downsample, use_softmax)

    def forward(self, x, segSize=None):
        if x.size(1) == 1:
            x = self.conv(x)
        _, x2, x3, x4 = self.backbone(x)
        return self.forward_head(x2, x3, x4)


class _V1Mixin(_PraNetBase):
    def forward_head(self, x2, x3, x4):
        """PraNet_Res2Net.py:143-186 -> (l5, l4, l3, l2)."""
        up = ops.interpolate_bilinear
        x2_rfb, x3_rfb, x4_rfb = self.rfb2_1(x2), self.rfb3_1(x3), self.rfb4_1(x4)
        ra5 = self.agg1(x4_rfb, x3_rfb, x2_rfb)
        l5 = up(ra5, scale_factor=8)
        crop = up(ra5, scale_factor=0.25)
        x = self.ra4_conv5(self._stack(4, ops.ra_v1_scale(x4, crop), 4)) + crop
        l4 = up(x, scale_factor=32)
        crop = up(x, scale_factor=2)
        x = self.ra3_conv4(self._stack(3, ops.ra_v1_scale(x3, crop), 3)) + crop
        l3 = up(x, scale_factor=16)
        crop = up(x, scale_factor=2)
        x = self.ra2_conv4(self._stack(2, ops.ra_v1_scale(x2, crop), 3)) + crop
        l2 = up(x, scale_factor=8)
        return l5, l4, l3, l2


class PraNet(_V1Mixin):
    """PraNet-V1 on Res2Net-50 (PraNet_Res2Net.py:101-186); backbone attribute is `resnet` there."""

    def __init__(self, channel=32):
        super().__init__()
        self.resnet = Res2Net50()
        self._build_head(RES2NET_CH, channel, 1, v1=True)

    def forward(self, x):
        _, x2, x3, x4 = self.resnet.pyramid(x)
        return self.forward_head(x2, x3, x4)


class PVT_PraNet(_V1Mixin):
    """PraNet-V1 on PVTv2-b2 (PraNet_Res2Net.py:188-270)."""

    def __init__(self, channel=32):
        super().__init__()
        self.backbone = PvtV2B2()
        self._build_head(PVT_CH, channel, 1, v1=True)

    def forward(self, x):
        _, x2, x3, x4 = self.backbone(x)
        return self.forward_head(x2, x3, x4)
