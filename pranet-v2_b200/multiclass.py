"""Drop-in carriers of the DSRA head in the multiclass integrations: the reference's dual-supervised decoders with the same
class names, constructor signatures, forward signatures / return tuples and state_dict keys, so checkpoints and training
loops written against the reference run unchanged:

    EMCAD_dual        multiclass_seg/EMCAD/lib/decoders.py:407-526   (host network EMCADNet: lib/networks.py:18-145)
    CASCADE_Add_dual  multiclass_seg/MERIT/lib/decoders.py:289-431   (returns the 9-tuple that ends with d1)
    CAM               multiclass_seg/MIST/lib/MIST.py:368-451        (dual mode: channels=..., n_class=...)

What is the hot path and what is context.  In all three decoders the DSRA part never feeds back into the decoder trunk: the
trunk features d4..d1 depend on the encoder pyramid only, and the DSRA stages read them.  So `forward` runs the trunk (the
host decoders' context blocks -- attention gates, channel / spatial attention, depth-wise stacks, the MIST transformer blocks)
as STOCK PyTorch modules, restated here compactly with the reference's parameter names, and hands [d4, d3, d2, d1] to
`heads.DSRAStages`: the fg / bg heads of a stage are ONE tcgen05 GEMM (N = 2 * num_class), BatchNorm statistics come out of the
GEMM launch, and the cascaded fusion  fg_k <- fg_k + fg_k * softmax_c(resize(fg_{k+1}) - resize(bg_{k+1}))  with its two bilinear
resizes is one kernel per stage (pv2_dsra_fuse_*).  SURVEY.md section 8 rows a7, a8, a13.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .backbones import PvtV2B2
from .heads import DSRAStages


def _normal_init(module: nn.Module):
    """The 'normal' scheme the EMCAD blocks apply to themselves (decoders.py:31-60): conv weights N(0, 0.02), biases 0, BN (1, 0)."""
    for m in module.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.BatchNorm2d):
            nn.init.ones_(m.weight)
            nn.init.zeros_(m.bias)


def _act(name: str) -> nn.Module:
    table = {"relu": nn.ReLU, "relu6": nn.ReLU6, "gelu": nn.GELU, "hswish": nn.Hardswish}
    if name.lower() not in table:
        raise NotImplementedError(f"activation layer [{name}] is not found")
    return table[name.lower()]()


def _shuffle(x, groups):
    b, c, h, w = x.shape
    return x.view(b, groups, c // groups, h, w).transpose(1, 2).reshape(b, c, h, w)


# ----------------------------------------------------------------------------------------------------------------------
# EMCAD context blocks (decoders.py:92-330), stock PyTorch
# ----------------------------------------------------------------------------------------------------------------------
class MSDC(nn.Module):
    def __init__(self, in_channels, kernel_sizes, stride, activation="relu6", dw_parallel=True):
        super().__init__()
        self.dw_parallel = dw_parallel
        self.dwconvs = nn.ModuleList([
            nn.Sequential(nn.Conv2d(in_channels, in_channels, k, stride, k // 2, groups=in_channels, bias=False),
                          nn.BatchNorm2d(in_channels), _act(activation)) for k in kernel_sizes])
        _normal_init(self)

    def forward(self, x):
        outs = []
        for dw in self.dwconvs:
            y = dw(x)
            outs.append(y)
            if not self.dw_parallel:
                x = x + y
        return outs


class MSCB(nn.Module):
    def __init__(self, in_channels, out_channels, stride, kernel_sizes=(1, 3, 5), expansion_factor=2, dw_parallel=True, add=True, activation="relu6"):
        super().__init__()
        assert stride in (1, 2)
        self.in_channels, self.out_channels, self.add = in_channels, out_channels, add
        self.use_skip_connection = stride == 1
        ex = int(in_channels * expansion_factor)
        self.pconv1 = nn.Sequential(nn.Conv2d(in_channels, ex, 1, 1, 0, bias=False), nn.BatchNorm2d(ex), _act(activation))
        self.msdc = MSDC(ex, kernel_sizes, stride, activation, dw_parallel=dw_parallel)
        self.combined_channels = ex if add else ex * len(kernel_sizes)
        self.pconv2 = nn.Sequential(nn.Conv2d(self.combined_channels, out_channels, 1, 1, 0, bias=False), nn.BatchNorm2d(out_channels))
        if self.use_skip_connection and in_channels != out_channels:
            self.conv1x1 = nn.Conv2d(in_channels, out_channels, 1, 1, 0, bias=False)
        _normal_init(self)

    def forward(self, x):
        outs = self.msdc(self.pconv1(x))
        d = sum(outs) if self.add else torch.cat(outs, dim=1)
        out = self.pconv2(_shuffle(d, math.gcd(self.combined_channels, self.out_channels)))
        if not self.use_skip_connection:
            return out
        return (self.conv1x1(x) if self.in_channels != self.out_channels else x) + out


def MSCBLayer(in_channels, out_channels, n=1, stride=1, **kw):
    return nn.Sequential(MSCB(in_channels, out_channels, stride, **kw), *[MSCB(out_channels, out_channels, 1, **kw) for _ in range(1, n)])


class EUCB(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, activation="relu"):
        super().__init__()
        self.in_channels = in_channels
        self.up_dwc = nn.Sequential(nn.Upsample(scale_factor=2),
                                    nn.Conv2d(in_channels, in_channels, kernel_size, stride, kernel_size // 2, groups=in_channels, bias=False),
                                    nn.BatchNorm2d(in_channels), _act(activation))
        self.pwc = nn.Sequential(nn.Conv2d(in_channels, out_channels, 1, 1, 0, bias=True))
        _normal_init(self)

    def forward(self, x):
        return self.pwc(_shuffle(self.up_dwc(x), self.in_channels))


class LGAG(nn.Module):
    def __init__(self, F_g, F_l, F_int, kernel_size=3, groups=1, activation="relu"):
        super().__init__()
        if kernel_size == 1:
            groups = 1
        self.W_g = nn.Sequential(nn.Conv2d(F_g, F_int, kernel_size, 1, kernel_size // 2, groups=groups, bias=True), nn.BatchNorm2d(F_int))
        self.W_x = nn.Sequential(nn.Conv2d(F_l, F_int, kernel_size, 1, kernel_size // 2, groups=groups, bias=True), nn.BatchNorm2d(F_int))
        self.psi = nn.Sequential(nn.Conv2d(F_int, 1, 1, 1, 0, bias=True), nn.BatchNorm2d(1), nn.Sigmoid())
        self.activation = _act(activation)
        _normal_init(self)

    def forward(self, g, x):
        return x * self.psi(self.activation(self.W_g(g) + self.W_x(x)))


class CAB(nn.Module):
    def __init__(self, in_channels, out_channels=None, ratio=16, activation="relu"):
        super().__init__()
        ratio = min(ratio, in_channels)
        self.activation = _act(activation)
        self.fc1 = nn.Conv2d(in_channels, in_channels // ratio, 1, bias=False)
        self.fc2 = nn.Conv2d(in_channels // ratio, out_channels or in_channels, 1, bias=False)
        _normal_init(self)

    def forward(self, x):
        f = lambda p: self.fc2(self.activation(self.fc1(p)))
        return torch.sigmoid(f(F.adaptive_avg_pool2d(x, 1)) + f(F.adaptive_max_pool2d(x, 1)))


class SAB(nn.Module):
    def __init__(self, kernel_size=7):
        super().__init__()
        assert kernel_size in (3, 7, 11), "kernel must be 3 or 7 or 11"
        self.conv = nn.Conv2d(2, 1, kernel_size, padding=kernel_size // 2, bias=False)
        _normal_init(self)

    def forward(self, x):
        return torch.sigmoid(self.conv(torch.cat([x.mean(1, keepdim=True), x.amax(1, keepdim=True)], dim=1)))


class EMCAD_dual(nn.Module):
    """EMCAD decoder with the DSRA stages (EMCAD/lib/decoders.py:407-526).  forward(x, skips) -> [d4_fg, d3_fg, d2_fg, d1_fg,
    d4_bg, d3_bg, d2_bg, d1_bg], deep -> shallow, each (B, num_class, h_k, w_k)."""

    def __init__(self, channels=[512, 320, 128, 64], kernel_sizes=[1, 3, 5], expansion_factor=6, dw_parallel=True, add=True, lgag_ks=3,
                 activation="relu6", num_class=None):
        super().__init__()
        assert num_class is not None
        kw = dict(kernel_sizes=kernel_sizes, expansion_factor=expansion_factor, dw_parallel=dw_parallel, add=add, activation=activation)
        c = channels
        self.mscb4 = MSCBLayer(c[0], c[0], n=1, stride=1, **kw)
        for k in (3, 2, 1):
            ci, co = c[3 - k], c[4 - k]
            setattr(self, f"eucb{k}", EUCB(ci, co, kernel_size=3, stride=1))
            setattr(self, f"lgag{k}", LGAG(F_g=co, F_l=co, F_int=co // 2, kernel_size=lgag_ks, groups=co // 2))
            setattr(self, f"mscb{k}", MSCBLayer(co, co, n=1, stride=1, **kw))
        for k in (4, 3, 2, 1):
            setattr(self, f"cab{k}", CAB(c[4 - k]))
        self.sab = SAB()
        # ConvBlock{4..1}_{fg,bg}: 1x1 for the deepest stage, 3x3 for the others, conv + BN (decoders.py:434-444)
        self._dsra = DSRAStages(self, c, num_class)

    def trunk(self, x, skips):
        d = self.cab4(x) * x
        d = self.sab(d) * d
        feats = [self.mscb4(d)]
        for k, skip in zip((3, 2, 1), skips):
            d = getattr(self, f"eucb{k}")(feats[-1])
            d = d + getattr(self, f"lgag{k}")(g=d, x=skip)
            d = getattr(self, f"cab{k}")(d) * d
            d = self.sab(d) * d
            feats.append(getattr(self, f"mscb{k}")(d))
        return feats

    def forward(self, x, skips):
        return self._dsra(self.trunk(x, skips))


class EMCADNet(nn.Module):
    """EMCADNet(encoder='pvt_v2_b2', dual=True) (EMCAD/lib/networks.py:18-145): grayscale stem, PVTv2-b2 pyramid, EMCAD_dual, the
    eight final upsamples x32 / x16 / x8 / x4 (networks.py:114-125).  `forward(x, mode)` returns the list the reference's trainer
    consumes (EMCAD/trainer.py:105-140); `forward_lowres` stops before the final upsamples."""

    def __init__(self, num_classes=1, kernel_sizes=[1, 3, 5], expansion_factor=2, dw_parallel=True, add=True, lgag_ks=3, activation="relu",
                 encoder="pvt_v2_b2", pretrain=False, dual=True):
        super().__init__()
        if encoder != "pvt_v2_b2" or not dual:
            raise NotImplementedError("pranet_v2_b200.EMCADNet covers the dual-supervised PVTv2-b2 configuration (the DSRA integration)")
        self.dual = True
        self.conv = nn.Sequential(nn.Conv2d(1, 3, kernel_size=1), nn.BatchNorm2d(3), nn.ReLU(inplace=True))
        self.backbone = PvtV2B2()
        self.decoder = EMCAD_dual(channels=[512, 320, 128, 64], kernel_sizes=kernel_sizes, expansion_factor=expansion_factor, dw_parallel=dw_parallel,
                                  add=add, lgag_ks=lgag_ks, activation=activation, num_class=num_classes)
        # single-supervision heads: defined (and checkpointed) by the reference in dual mode too, never applied there (networks.py:90-93)
        for k, c in zip((4, 3, 2, 1), (512, 320, 128, 64)):
            setattr(self, f"out_head{k}", nn.Conv2d(c, num_classes, 1))
        self.scale_factors = (32, 16, 8, 4)

    def forward_lowres(self, x):
        if x.size(1) == 1:
            x = self.conv(x)
        x1, x2, x3, x4 = self.backbone(x)
        return self.decoder(x4, [x3, x2, x1])

    def forward(self, x, mode="test"):
        outs = self.forward_lowres(x)
        return [ops.interpolate_bilinear(o.float(), scale_factor=float(s)) for o, s in zip(outs, self.scale_factors * 2)]


# ----------------------------------------------------------------------------------------------------------------------
# MERIT context blocks (MERIT/lib/decoders.py:19-120) and the dual cascade decoder
# ----------------------------------------------------------------------------------------------------------------------
class conv_block(nn.Module):
    def __init__(self, ch_in, ch_out):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(ch_in, ch_out, 3, 1, 1, bias=True), nn.BatchNorm2d(ch_out), nn.ReLU(inplace=True),
                                  nn.Conv2d(ch_out, ch_out, 3, 1, 1, bias=True), nn.BatchNorm2d(ch_out), nn.ReLU(inplace=True))

    def forward(self, x):
        return self.conv(x)


class up_conv(nn.Module):
    def __init__(self, ch_in, ch_out, kernel_size=3, stride=1, padding=1, groups=1):
        super().__init__()
        self.up = nn.Sequential(nn.Upsample(scale_factor=2), nn.Conv2d(ch_in, ch_out, kernel_size, stride, padding, bias=True),
                                nn.BatchNorm2d(ch_out), nn.ReLU(inplace=True))

    def forward(self, x):
        return self.up(x)


class Attention_block(nn.Module):
    def __init__(self, F_g, F_l, F_int):
        super().__init__()
        self.W_g = nn.Sequential(nn.Conv2d(F_g, F_int, 1, 1, 0, bias=True), nn.BatchNorm2d(F_int))
        self.W_x = nn.Sequential(nn.Conv2d(F_l, F_int, 1, 1, 0, bias=True), nn.BatchNorm2d(F_int))
        self.psi = nn.Sequential(nn.Conv2d(F_int, 1, 1, 1, 0, bias=True), nn.BatchNorm2d(1), nn.Sigmoid())
        self.relu = nn.ReLU(inplace=True)

    def forward(self, g, x):
        return x * self.psi(self.relu(self.W_g(g) + self.W_x(x)))


class ChannelAttention(nn.Module):
    def __init__(self, in_planes, ratio=16):
        super().__init__()
        self.fc1 = nn.Conv2d(in_planes, in_planes // 16, 1, bias=False)
        self.relu1 = nn.ReLU()
        self.fc2 = nn.Conv2d(in_planes // 16, in_planes, 1, bias=False)

    def forward(self, x):
        f = lambda p: self.fc2(self.relu1(self.fc1(p)))
        return torch.sigmoid(f(F.adaptive_avg_pool2d(x, 1)) + f(F.adaptive_max_pool2d(x, 1)))


class SpatialAttention(nn.Module):
    def __init__(self, kernel_size=7):
        super().__init__()
        assert kernel_size in (3, 7), "kernel size must be 3 or 7"
        self.conv1 = nn.Conv2d(2, 1, kernel_size, padding=3 if kernel_size == 7 else 1, bias=False)

    def forward(self, x):
        return torch.sigmoid(self.conv1(torch.cat([x.mean(1, keepdim=True), x.amax(1, keepdim=True)], dim=1)))


class CASCADE_Add_dual(nn.Module):
    """MERIT's additive cascade decoder with the DSRA stages (MERIT/lib/decoders.py:289-431).  forward(x, skips) ->
    (d4_fg, d3_fg, d2_fg, d1_fg, d4_bg, d3_bg, d2_bg, d1_bg, d1): the 9-tuple, d1 being the shallowest trunk feature."""

    def __init__(self, channels=[512, 320, 128, 64], num_class=None, use_softmax=True):
        super().__init__()
        assert num_class is not None
        self.use_softmax = use_softmax
        c = channels
        self.Conv_1x1 = nn.Conv2d(c[0], c[0], kernel_size=1, stride=1, padding=0)
        self.ConvBlock4 = conv_block(c[0], c[0])
        f_int = (c[2], c[3], int(c[3] / 2))
        for i, k in enumerate((3, 2, 1)):
            ci, co = c[i], c[i + 1]
            setattr(self, f"Up{k}", up_conv(ci, co))
            setattr(self, f"AG{k}", Attention_block(F_g=co, F_l=co, F_int=f_int[i]))
            setattr(self, f"ConvBlock{k}", conv_block(co, co))
        for k in (4, 3, 2, 1):
            setattr(self, f"CA{k}", ChannelAttention(c[4 - k]))
        self.SA = SpatialAttention()
        self._dsra = DSRAStages(self, c, num_class, use_softmax=use_softmax)

    def trunk(self, x, skips):
        d = self.Conv_1x1(x)
        d = self.CA4(d) * d
        d = self.SA(d) * d
        feats = [self.ConvBlock4(d)]
        for k, skip in zip((3, 2, 1), skips):
            d = getattr(self, f"Up{k}")(feats[-1])
            d = d + getattr(self, f"AG{k}")(g=d, x=skip)
            d = getattr(self, f"CA{k}")(d) * d
            d = self.SA(d) * d
            feats.append(getattr(self, f"ConvBlock{k}")(d))
        return feats

    def forward(self, x, skips):
        feats = self.trunk(x, skips)
        return (*self._dsra(feats), feats[-1])


# ----------------------------------------------------------------------------------------------------------------------
# MIST context blocks (MIST/lib/MIST.py:24-100, 169-272, 327-366) and the dual CAM decoder
# ----------------------------------------------------------------------------------------------------------------------
def _ln_nchw(ln: nn.LayerNorm, x):
    return ln(x.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)


class Attention(nn.Module):
    """Convolutional-projection multi-head attention (MIST.py:24-100): depth-wise 3x3 + ReLU + LayerNorm for q, k, v."""

    def __init__(self, channels, num_heads, proj_drop=0.0, kernel_size=3, stride_kv=1, stride_q=1, padding_kv="same", padding_q="same",
                 attention_bias=True):
        super().__init__()
        self.proj_drop = proj_drop
        self.conv_q = nn.Conv2d(channels, channels, kernel_size, stride_q, padding_q, bias=attention_bias, groups=channels)
        self.layernorm_q = nn.LayerNorm(channels, eps=1e-5)
        self.conv_k = nn.Conv2d(channels, channels, kernel_size, stride_kv, stride_kv, bias=attention_bias, groups=channels)
        self.layernorm_k = nn.LayerNorm(channels, eps=1e-5)
        self.conv_v = nn.Conv2d(channels, channels, kernel_size, stride_kv, stride_kv, bias=attention_bias, groups=channels)
        self.layernorm_v = nn.LayerNorm(channels, eps=1e-5)
        self.attention = nn.MultiheadAttention(embed_dim=channels, bias=attention_bias, batch_first=True, num_heads=num_heads)

    def forward(self, x):
        b, c, h, w = x.shape
        proj = lambda conv, ln: _ln_nchw(ln, F.relu(conv(x))).reshape(b, c, h * w).permute(0, 2, 1)
        q, k, v = proj(self.conv_q, self.layernorm_q), proj(self.conv_k, self.layernorm_k), proj(self.conv_v, self.layernorm_v)
        y = self.attention(query=q, value=v, key=k, need_weights=False)[0].permute(0, 2, 1)
        side = int(np.sqrt(y.shape[2]))
        return F.dropout(y.reshape(b, c, side, side), self.proj_drop)


class Dilated_Conv(nn.Module):
    """Wide-focus block (MIST.py:214-243); the functional dropouts are the reference's (active in eval too)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, 1, padding="same")
        self.conv2 = nn.Conv2d(in_channels, out_channels, 3, 1, padding="same", dilation=2)
        self.conv3 = nn.Conv2d(in_channels, out_channels, 3, 1, padding="same", dilation=3)
        self.conv4 = nn.Conv2d(in_channels, out_channels, 3, 1, padding="same")

    def forward(self, x):
        branch = lambda conv: F.dropout(F.gelu(conv(x)), 0.1)
        added = branch(self.conv1) + branch(self.conv2) + branch(self.conv3)
        return F.dropout(F.gelu(self.conv4(added)), 0.1)


class Transformer(nn.Module):
    def __init__(self, out_channels, num_heads, dpr, proj_drop=0.0, attention_bias=True, padding_q="same", padding_kv="same", stride_kv=1, stride_q=1):
        super().__init__()
        self.attention_output = Attention(channels=out_channels, num_heads=num_heads, proj_drop=proj_drop, padding_q=padding_q, padding_kv=padding_kv,
                                          stride_kv=stride_kv, stride_q=stride_q, attention_bias=attention_bias)
        self.conv1 = nn.Conv2d(out_channels, out_channels, 3, 1, padding="same")
        self.layernorm = nn.LayerNorm(out_channels, eps=1e-5)
        self.wide_focus = Dilated_Conv(out_channels, out_channels)

    def forward(self, x):
        x2 = self.conv1(self.attention_output(x)) + x
        return x2 + self.wide_focus(_ln_nchw(self.layernorm, x2))


class Block_decoder(nn.Module):
    def __init__(self, in_channels, out_channels, att_heads, dpr):
        super().__init__()
        self.layernorm = nn.LayerNorm(in_channels, eps=1e-5)
        self.upsample = nn.Upsample(scale_factor=2)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, 1, padding="same")
        self.conv2 = nn.Conv2d(out_channels * 2, out_channels, 3, 1, padding="same")
        self.conv3 = nn.Conv2d(out_channels, out_channels, 3, 1, padding="same")      # defined (and checkpointed) by the reference, unused
        self.trans = Transformer(out_channels, att_heads, dpr)

    def forward(self, x, skip):
        x1 = F.relu(self.conv1(self.upsample(_ln_nchw(self.layernorm, x))))
        x1 = F.dropout(F.relu(self.conv2(torch.cat((skip, x1), dim=1))), 0.3)
        return self.trans(x1)


class Block_encoder_bottleneck(nn.Module):
    def __init__(self, blk, in_channels, out_channels, att_heads, dpr):
        super().__init__()
        if blk not in ("first", "bottleneck"):
            raise NotImplementedError("CAM uses the bottleneck form of this block only")
        self.blk = blk
        self.layernorm = nn.LayerNorm(in_channels, eps=1e-5)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, 1, padding="same")
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, 1, padding="same")
        self.trans = Transformer(out_channels, att_heads, dpr)

    def forward(self, x, scale_img="none"):
        x1 = F.relu(self.conv2(F.relu(self.conv1(_ln_nchw(self.layernorm, x)))))
        return self.trans(F.max_pool2d(F.dropout(x1, 0.3), (2, 2)))


class CAM(nn.Module):
    """MIST's convolutional-attention-mixing decoder in dual mode (MIST/lib/MIST.py:368-451): CAM(args, channels=[768, 384, 192, 96],
    n_class=9).forward(skip1, skip2, skip3, skip4) -> (d4_fg, d3_fg, d2_fg, d1_fg, d4_bg, d3_bg, d2_bg, d1_bg).  The DSRA heads
    `out_head{1..4}_{fg,bg}` are biased 1x1 convolutions without BatchNorm (:403-412)."""

    def __init__(self, args=None, **kwargs):
        super().__init__()
        att_heads = [2, 4, 8, 12, 16, 12, 8, 4, 2]
        filters = [96, 192, 384, 768, 768 * 2, 768, 384, 192, 96]
        dpr = list(np.linspace(0, 1.0, len(filters)))
        self.drp_out = 0.3
        self.scale_img = nn.AvgPool2d(2, 2)
        self.block_5 = Block_encoder_bottleneck("bottleneck", filters[3], filters[4], att_heads[4], dpr[4])
        for i in (6, 7, 8, 9):
            setattr(self, f"block_{i}", Block_decoder(filters[i - 2], filters[i - 1], att_heads[i - 1], dpr[i - 1]))
        self.channels = kwargs.get("channels", False)
        self.n_class = kwargs.get("n_class", 0)
        if not (self.channels and self.n_class != 0):
            raise NotImplementedError("pranet_v2_b200.CAM is the dual-supervised (DSRA) configuration: pass channels=[...] and n_class")
        self._dsra = DSRAStages(self, self.channels, self.n_class, names=("out_head1", "out_head2", "out_head3", "out_head4"),
                                kernel_sizes=(1, 1, 1, 1), bn=False)

    def trunk(self, skip1, skip2, skip3, skip4):
        x = self.block_5(skip4)
        feats = []
        for blk, skip in zip((self.block_6, self.block_7, self.block_8, self.block_9), (skip4, skip3, skip2, skip1)):
            x = blk(x, skip)
            feats.append(x)
        return feats

    def forward(self, skip1, skip2, skip3, skip4):
        return tuple(self._dsra(self.trunk(skip1, skip2, skip3, skip4)))
