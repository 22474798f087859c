"""state_dict key/shape tables of the reference's head modules, restated by hand.

TEST INFRASTRUCTURE.  Lets the oracle (and the key-compatibility tests) know the 425 head
keys of PraNet_V2 without importing the reference: BasicConv2d -> conv.weight + bn.{weight,
bias,running_mean,running_var,num_batches_tracked} (binary_seg/lib/pranet.py:31-37);
RFB_modified (pranet.py:51-73); aggregation (pranet.py:94-104, V1: PraNet_Res2Net.py:72-81);
DSRA stacks (pranet.py:303-325, V1: PraNet_Res2Net.py:114-128).
"""
from __future__ import annotations

import torch


def _pair(k):
    return (k, k) if isinstance(k, int) else tuple(k)


def basic_conv(prefix, cin, cout, k):
    kh, kw = _pair(k)
    z = torch.zeros
    return {
        f"{prefix}.conv.weight": z(cout, cin, kh, kw),
        f"{prefix}.bn.weight": z(cout), f"{prefix}.bn.bias": z(cout),
        f"{prefix}.bn.running_mean": z(cout), f"{prefix}.bn.running_var": z(cout),
        f"{prefix}.bn.num_batches_tracked": z((), dtype=torch.long),
    }


def rfb(prefix, cin, c):
    d = {}
    d.update(basic_conv(f"{prefix}.branch0.0", cin, c, 1))
    for b, k in ((1, 3), (2, 5), (3, 7)):
        d.update(basic_conv(f"{prefix}.branch{b}.0", cin, c, 1))
        d.update(basic_conv(f"{prefix}.branch{b}.1", c, c, (1, k)))
        d.update(basic_conv(f"{prefix}.branch{b}.2", c, c, (k, 1)))
        d.update(basic_conv(f"{prefix}.branch{b}.3", c, c, 3))
    d.update(basic_conv(f"{prefix}.conv_cat", 4 * c, c, 3))
    d.update(basic_conv(f"{prefix}.conv_res", cin, c, 1))
    return d


def aggregation(prefix, c, num_class=None):
    d = {}
    for i in (1, 2, 3, 4):
        d.update(basic_conv(f"{prefix}.conv_upsample{i}", c, c, 3))
    d.update(basic_conv(f"{prefix}.conv_upsample5", 2 * c, 2 * c, 3))
    d.update(basic_conv(f"{prefix}.conv_concat2", 2 * c, 2 * c, 3))
    d.update(basic_conv(f"{prefix}.conv_concat3", 3 * c, 3 * c, 3))
    d.update(basic_conv(f"{prefix}.conv4", 3 * c, 3 * c, 3))
    heads = ("conv5",) if num_class is None else ("conv5_fg", "conv5_bg")
    for h in heads:
        d[f"{prefix}.{h}.weight"] = torch.zeros(num_class or 1, 3 * c, 1, 1)
        d[f"{prefix}.{h}.bias"] = torch.zeros(num_class or 1)
    return d


def pranet_head(channels=(512, 1024, 2048), channel=32, num_class=1, v1=False):
    c2, c3, c4 = channels
    d = {}
    d.update(rfb("rfb2_1", c2, channel))
    d.update(rfb("rfb3_1", c3, channel))
    d.update(rfb("rfb4_1", c4, channel))
    d.update(aggregation("agg1", channel, None if v1 else num_class))
    d.update(basic_conv("ra4_conv1", c4, 256, 1))
    for i in (2, 3, 4):
        d.update(basic_conv(f"ra4_conv{i}", 256, 256, 5))
    for s, cin in ((3, c3), (2, c2)):
        d.update(basic_conv(f"ra{s}_conv1", cin, 64, 1))
        d.update(basic_conv(f"ra{s}_conv2", 64, 64, 3))
        d.update(basic_conv(f"ra{s}_conv3", 64, 64, 3))
    if v1:
        d.update(basic_conv("ra4_conv5", 256, 1, 1))
        d.update(basic_conv("ra3_conv4", 64, 1, 3))
        d.update(basic_conv("ra2_conv4", 64, 1, 3))
    else:
        for t in ("fg", "bg"):
            d.update(basic_conv(f"ra4_conv5_{t}", 256, num_class, 1))
            d.update(basic_conv(f"ra3_conv4_{t}", 64, num_class, 3))
            d.update(basic_conv(f"ra2_conv4_{t}", 64, num_class, 3))
    return d


def dual_heads(channels, num_class, names=("ConvBlock4", "ConvBlock3", "ConvBlock2", "ConvBlock1"),
               kernel_sizes=(1, 3, 3, 3), bn=True):
    """DSRA head keys of the multiclass carriers (EMCAD/lib/decoders.py:434-444,
    MERIT/lib/decoders.py:298-322; bn=False: MIST/lib/MIST.py:403-412)."""
    d = {}
    for c, n, k in zip(channels, names, kernel_sizes):
        for t in ("fg", "bg"):
            if bn:
                d.update(basic_conv(f"{n}_{t}", c, num_class, k))
            else:
                d[f"{n}_{t}.weight"] = torch.zeros(num_class, c, 1, 1)
                d[f"{n}_{t}.bias"] = torch.zeros(num_class)
    return d
