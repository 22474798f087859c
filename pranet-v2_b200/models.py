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

from . import engine as E
from . import ops
from .backbones import PvtV2B2, Res2Net50
from .heads import (BasicConv2d, RFB_modified, _sink, aggregation, aggregation_run, rfb_begin, rfb_bns, rfb_branch, rfb_convs, rfb_finish,
                    rfb_run)

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

    def _stack(self, eng, stage, t, n):
        """ra{stage}_conv2..n, each followed by ReLU (pranet.py:358-360, 379-380, 401-402); `t` is the BN'd (no ReLU)
        output of ra{stage}_conv1."""
        for i in range(2, n + 1):
            t = getattr(self, f"ra{stage}_conv{i}").run(eng, t, relu=True)
        return t

    def head_parameters(self):
        return [p for n, p in self.named_parameters() if not n.startswith(("backbone.", "resnet.", "conv."))]

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

    def final_scale_factors(self):
        """scale_factor of the four final F.interpolate calls (pranet.py:349-350,370-371,392-393,414-415), in output order."""
        sd = self.sem_downsample
        return [8 / sd, 16 / sd, 32 / sd, 8 / sd]

    def forward_head(self, x2, x3, x4, lowres=False):
        """pranet.py:343-417 -> (l2_fg, l3_fg, l4_fg, l5_fg, l2_bg, l3_bg, l4_bg, l5_bg).
        lowres=True returns the same eight maps BEFORE their final upsample (44^2, 22^2, 11^2, 44^2 at 352^2 input): what the
        fused inference tail (ops.infer_tail_binary) and loss-from-low-res paths consume."""
        sd = self.sem_downsample

        def runner(eng, inputs, in_grads):
            feats = [eng.from_nchw(t, _sink(in_grads, i)) for i, t in enumerate(inputs)]
            levels = list(zip(feats, (self.rfb2_1, self.rfb3_1, self.rfb4_1), (2, 3, 4), (3, 3, 4)))
            # section 1 -- per pyramid level ONE GEMM for the six 1x1 convs that read the backbone feature:
            # [rfb.branch0..3 first convs | rfb.conv_res | ra_conv1]; the three levels are independent
            raws = [None] * 3
            eng.fork(3)
            for li, (a, rfb, stage, _) in enumerate(levels):
                with eng.branch(li + 1):
                    ra1 = getattr(self, f"ra{stage}_conv1")
                    raws[li] = eng.conv(a, rfb_convs(rfb) + [ra1.conv], rfb_bns(rfb) + [ra1.bn])
            eng.join()
            # section 2 -- per level: RFB branches 1..3, and [RFB branch 0 + the reverse-attention stack + its fg/bg heads]:
            # twelve independent chains
            states = [rfb_begin(eng, rfb, raws[li]) for li, (_, rfb, _, _) in enumerate(levels)]
            heads_ = [None] * 3
            eng.fork(12)
            for li, (a, rfb, stage, depth) in enumerate(levels):
                for b in (1, 2, 3):
                    with eng.branch(li * 4 + b):
                        rfb_branch(eng, rfb, states[li], raws[li], 0, b)
                with eng.branch(li * 4 + 4):
                    rfb_branch(eng, rfb, states[li], raws[li], 0, 0)
                    ra1 = getattr(self, f"ra{stage}_conv1")
                    c = rfb.conv_res.conv.out_channels
                    t = eng.bn_apply((raws[li], 5 * c, ra1.conv.out_channels, ra1.bn, None))                 # no ReLU after conv1
                    t = self._stack(eng, stage, t, depth)
                    head = "conv5" if stage == 4 else "conv4"
                    heads_[li] = dual(eng, getattr(self, f"ra{stage}_{head}_fg"), getattr(self, f"ra{stage}_{head}_bg"), t)
            eng.join()
            # section 3 -- conv_cat + residual + ReLU of the three RFBs
            rfbs = [None] * 3
            eng.fork(3)
            for li, (_, rfb, _, _) in enumerate(levels):
                with eng.branch(li + 1):
                    rfbs[li] = rfb_finish(eng, rfb, states[li], raws[li], 0)
            eng.join()
            ra5_fg, ra5_bg = aggregation_run(eng, self.agg1, rfbs[2], rfbs[1], rfbs[0])
            (fg2, bg2), (fg3, bg3), (fg4, bg4) = heads_
            # DSRA fusions, deep -> shallow; the x0.25 / x2 resize of the deeper maps is fused into the fusion kernel
            fg4 = eng.fuse(fg4, ra5_fg, ra5_bg, self.use_softmax, 0.25)
            fg3 = eng.fuse(fg3, fg4, bg4, self.use_softmax, 2)
            fg2 = eng.fuse(fg2, fg3, bg3, self.use_softmax, 2)
            if lowres:
                return [fg2, fg3, fg4, ra5_fg, bg2, bg3, bg4, ra5_bg]
            # the eight final upsamples (x8, x16, x32, x8 for fg and bg; pranet.py:349-350,370-371,392-393,414-415): one launch
            l2_fg, l3_fg, l4_fg, l5_fg, l2_bg, l3_bg, l4_bg, l5_bg = eng.resize_multi(
                [fg2, fg3, fg4, ra5_fg, bg2, bg3, bg4, ra5_bg], [8 / sd, 16 / sd, 32 / sd, 8 / sd] * 2)
            return [l2_fg, l3_fg, l4_fg, l5_fg, l2_bg, l3_bg, l4_bg, l5_bg]

        from .heads import dual_heads_run as dual
        cache = self.__dict__.setdefault("_pv2_cache", {})
        cache["early_groups"] = 3          # the three level GEMMs open the head: their weights are packed first, the rest meanwhile
        return E.run_head(runner, [x2, x3, x4], self.head_parameters(), self.training, cache)


    @torch.no_grad()
    def predict_uint8(self, x, size=None):
        """The whole test-time path of binary_seg/MyTest_med.py:35-42 (:104-111) for a batch: forward, p2+p3+p4+p5, resize to
        `size` (the ground-truth H x W; None = input size), sigmoid, per-image min-max, uint8 -- the post-processing is ONE
        fused tail on the low-res maps (num_class must be 1, as in the reference's polyp models)."""
        if self.num_class != 1:
            raise ValueError("predict_uint8 is the binary test path (num_class=1); use ops.infer_tail_argmax for multiclass maps")
        maps = self.forward_features_lowres(x)[:4]
        return ops.infer_tail_binary(maps, self.final_scale_factors(), size)


class PraNet_V2(_V2Mixin):
    """Res2Net-50 PraNet-V2 (binary_seg/lib/pranet.py:268-417)."""

    def __init__(self, channel=32, num_class=3, sem_downsample=1, use_softmax=True):
        super().__init__()
        self.backbone = Res2Net50()
        self._init_v2(RES2NET_CH, channel, num_class, sem_downsample, use_softmax)

    def forward(self, x, segSize=None):
        _, x2, x3, x4 = self.backbone.pyramid(x)
        return self.forward_head(x2, x3, x4)

    def forward_features_lowres(self, x):
        _, x2, x3, x4 = self.backbone.pyramid(x)
        return self.forward_head(x2, x3, x4, lowres=True)


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

    def forward_features_lowres(self, x):
        if x.size(1) == 1:
            x = self.conv(x)
        _, x2, x3, x4 = self.backbone(x)
        return self.forward_head(x2, x3, x4, lowres=True)


class _V1Mixin(_PraNetBase):
    def forward_head(self, x2, x3, x4):
        """PraNet_Res2Net.py:143-186 -> (l5, l4, l3, l2)."""
        def runner(eng, inputs, in_grads):
            feats = [eng.from_nchw(t, _sink(in_grads, i)) for i, t in enumerate(inputs)]
            rfbs = [rfb_run(eng, rfb, eng.conv(a, rfb_convs(rfb), rfb_bns(rfb)), 0) for a, rfb in zip(feats, (self.rfb2_1, self.rfb3_1, self.rfb4_1))]
            ra5 = aggregation_run(eng, self.agg1, rfbs[2], rfbs[1], rfbs[0])[0]
            outs = [eng.resize(ra5, 8)]
            x = ra5
            for stage, i, n, head, s_in, s_out in ((4, 2, 4, self.ra4_conv5, 0.25, 32), (3, 1, 3, self.ra3_conv4, 2, 16), (2, 0, 3, self.ra2_conv4, 2, 8)):
                crop = eng.resize(x, s_in)
                # reverse attention: (1 - sigmoid(crop)) * x_k  (PraNet_Res2Net.py:153-154), then the conv stack
                scaled, sink = eng.ra_v1(inputs[i], crop, _sink(in_grads, i))
                t = getattr(self, f"ra{stage}_conv1").run(eng, eng.from_nchw(scaled, sink))
                t = self._stack(eng, stage, t, n)
                x = eng.add_maps(head.run(eng, t, out_map=True), crop)                                  # ra_feat + crop
                outs.append(eng.resize(x, s_out))
            return outs
        return E.run_head(runner, [x2, x3, x4], self.head_parameters(), self.training, self.__dict__.setdefault("_pv2_cache", {}))


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
