"""Stock-PyTorch backbones (OUT OF SCOPE context for the DSRA hot path -- SURVEY.md §2 rows 4-5).

They exist so that the drop-in host networks can run end to end and so that the reference's
checkpoints load unchanged: parameter / buffer names follow the reference's state_dict layout
(`binary_seg/lib/Res2Net_v1b.py`: conv1.{0,1,3,4,6}, bn1, layer{1..4}.{i}.{conv1,bn1,convs.j,bns.j,
conv3,bn3,downsample.{1,2}}, fc;  `binary_seg/lib/pvtv2.py`: patch_embed{k}.{proj,norm},
block{k}.{i}.{norm1,attn.{q,kv,proj,sr,norm},norm2,mlp.{fc1,dwconv.dwconv,fc2}}, norm{k}).
Nothing here is a custom kernel: plain nn.Conv2d / nn.Linear / SDPA.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# Res2Net-50 v1b, 26w x 4s
# ----------------------------------------------------------------------------------------------
class _Res2Block(nn.Module):
    """Bottleneck whose 3x3 is split into `scale` channel groups processed hierarchically."""

    def __init__(self, cin, planes, stride=1, first_of_stage=False, base_width=26, scale=4, shortcut=None):
        super().__init__()
        self.width = width = int(math.floor(planes * base_width / 64.0))
        self.scale, self.first = scale, first_of_stage
        self.conv1 = nn.Conv2d(cin, width * scale, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(width * scale)
        self.convs = nn.ModuleList(nn.Conv2d(width, width, 3, stride, 1, bias=False) for _ in range(scale - 1))
        self.bns = nn.ModuleList(nn.BatchNorm2d(width) for _ in range(scale - 1))
        if first_of_stage:
            self.pool = nn.AvgPool2d(3, stride, 1)
        self.conv3 = nn.Conv2d(width * scale, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.downsample = shortcut

    def forward(self, x):
        y = F.relu(self.bn1(self.conv1(x)))
        groups = y.split(self.width, dim=1)
        outs, carry = [], None
        for j, (conv, bn) in enumerate(zip(self.convs, self.bns)):
            inp = groups[j] if (j == 0 or self.first) else carry + groups[j]
            carry = F.relu(bn(conv(inp)))
            outs.append(carry)
        outs.append(self.pool(groups[-1]) if self.first else groups[-1])
        y = self.bn3(self.conv3(torch.cat(outs, 1)))
        return F.relu(y + (x if self.downsample is None else self.downsample(x)))


class Res2Net50(nn.Module):
    """Res2Net-50 v1b 26w4s (deep 3-conv stem, avg-pool shortcuts)."""

    def __init__(self, depths=(3, 4, 6, 3), num_classes=1000):
        super().__init__()
        self.conv1 = nn.Sequential(
            nn.Conv2d(3, 32, 3, 2, 1, bias=False), nn.BatchNorm2d(32), nn.ReLU(inplace=True),
            nn.Conv2d(32, 32, 3, 1, 1, bias=False), nn.BatchNorm2d(32), nn.ReLU(inplace=True),
            nn.Conv2d(32, 64, 3, 1, 1, bias=False))
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU()
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        cin = 64
        for idx, (planes, n) in enumerate(zip((64, 128, 256, 512), depths), start=1):
            stride = 1 if idx == 1 else 2
            shortcut = nn.Sequential(
                nn.AvgPool2d(stride, stride, ceil_mode=True, count_include_pad=False),
                nn.Conv2d(cin, planes * 4, 1, bias=False), nn.BatchNorm2d(planes * 4))
            blocks = [_Res2Block(cin, planes, stride, True, shortcut=shortcut)]
            cin = planes * 4
            blocks += [_Res2Block(cin, planes) for _ in range(n - 1)]
            setattr(self, f"layer{idx}", nn.Sequential(*blocks))
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Linear(cin, num_classes)   # unused by the segmentation nets; kept for checkpoint keys
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")

    def pyramid(self, x):
        """-> (x1 /4 256ch, x2 /8 512ch, x3 /16 1024ch, x4 /32 2048ch)"""
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        x1 = self.layer1(x)
        x2 = self.layer2(x1)
        x3 = self.layer3(x2)
        x4 = self.layer4(x3)
        return x1, x2, x3, x4

    def forward(self, x):
        return self.fc(self.avgpool(self.pyramid(x)[-1]).flatten(1))


# ----------------------------------------------------------------------------------------------
# PVTv2-b2
# ----------------------------------------------------------------------------------------------
class _DropPath(nn.Module):
    def __init__(self, p=0.0):
        super().__init__()
        self.drop_prob = p

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        return x * mask / keep


class _DW(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.dwconv = nn.Conv2d(dim, dim, 3, 1, 1, groups=dim)

    def forward(self, x, H, W):
        B, N, Cc = x.shape
        return self.dwconv(x.transpose(1, 2).reshape(B, Cc, H, W)).flatten(2).transpose(1, 2)


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.dwconv = _DW(hidden)
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x, H, W):
        return self.fc2(F.gelu(self.dwconv(self.fc1(x), H, W)))


class _SRAttention(nn.Module):
    """Multi-head attention whose keys/values come from a spatially reduced copy of the tokens."""

    def __init__(self, dim, heads, sr_ratio):
        super().__init__()
        self.heads, self.sr_ratio = heads, sr_ratio
        self.q = nn.Linear(dim, dim)
        self.kv = nn.Linear(dim, dim * 2)
        self.proj = nn.Linear(dim, dim)
        if sr_ratio > 1:
            self.sr = nn.Conv2d(dim, dim, sr_ratio, sr_ratio)
            self.norm = nn.LayerNorm(dim)

    def forward(self, x, H, W):
        B, N, Cc = x.shape
        hd = Cc // self.heads
        q = self.q(x).view(B, N, self.heads, hd).transpose(1, 2)
        src = x
        if self.sr_ratio > 1:
            src = self.norm(self.sr(x.transpose(1, 2).reshape(B, Cc, H, W)).flatten(2).transpose(1, 2))
        kv = self.kv(src).view(B, -1, 2, self.heads, hd).permute(2, 0, 3, 1, 4)
        out = F.scaled_dot_product_attention(q, kv[0], kv[1])
        return self.proj(out.transpose(1, 2).reshape(B, N, Cc))


class _PvtBlock(nn.Module):
    def __init__(self, dim, heads, mlp_ratio, sr_ratio, drop_path):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _SRAttention(dim, heads, sr_ratio)
        self.drop_path = _DropPath(drop_path) if drop_path > 0 else nn.Identity()
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))

    def forward(self, x, H, W):
        x = x + self.drop_path(self.attn(self.norm1(x), H, W))
        return x + self.drop_path(self.mlp(self.norm2(x), H, W))


class _PatchEmbed(nn.Module):
    def __init__(self, k, stride, cin, dim):
        super().__init__()
        self.proj = nn.Conv2d(cin, dim, k, stride, k // 2)
        self.norm = nn.LayerNorm(dim)

    def forward(self, x):
        x = self.proj(x)
        H, W = x.shape[-2:]
        return self.norm(x.flatten(2).transpose(1, 2)), H, W


class PvtV2B2(nn.Module):
    """PVTv2-b2: dims 64/128/320/512, depths 3/4/6/3, heads 1/2/5/8, sr 8/4/2/1, drop-path 0.1."""

    def __init__(self, dims=(64, 128, 320, 512), depths=(3, 4, 6, 3), heads=(1, 2, 5, 8),
                 mlp_ratios=(8, 8, 4, 4), sr_ratios=(8, 4, 2, 1), drop_path_rate=0.1):
        super().__init__()
        rates = torch.linspace(0, drop_path_rate, sum(depths)).tolist()
        cin, cur = 3, 0
        for k in range(4):
            setattr(self, f"patch_embed{k + 1}", _PatchEmbed(7 if k == 0 else 3, 4 if k == 0 else 2, cin, dims[k]))
            setattr(self, f"block{k + 1}", nn.ModuleList(
                _PvtBlock(dims[k], heads[k], mlp_ratios[k], sr_ratios[k], rates[cur + i]) for i in range(depths[k])))
            setattr(self, f"norm{k + 1}", nn.LayerNorm(dims[k], eps=1e-6))
            cin, cur = dims[k], cur + depths[k]
        self.apply(self._init)

    @staticmethod
    def _init(m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=0.02)
            nn.init.zeros_(m.bias)
        elif isinstance(m, nn.Conv2d):
            fan_out = m.kernel_size[0] * m.kernel_size[1] * m.out_channels // m.groups
            nn.init.normal_(m.weight, 0, math.sqrt(2.0 / fan_out))
            if m.bias is not None:
                nn.init.zeros_(m.bias)

    def forward(self, x):
        B, outs = x.shape[0], []
        for k in range(1, 5):
            x, H, W = getattr(self, f"patch_embed{k}")(x)
            for blk in getattr(self, f"block{k}"):
                x = blk(x, H, W)
            x = getattr(self, f"norm{k}")(x).reshape(B, H, W, -1).permute(0, 3, 1, 2).contiguous()
            outs.append(x)
        return outs
