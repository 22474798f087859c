"""Dispatch layer between the drop-in modules (heads.py / models.py) and the kernels.

Every dense contraction, BN, activation and glue op of the head goes through one of the functions
below, so the choice of implementation is made in exactly one place.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops


def conv_bn_act(x, conv: nn.Conv2d, bn: nn.BatchNorm2d, relu: bool):
    """BasicConv2d body: BN(conv(x)) [+ ReLU] (binary_seg/lib/pranet.py:40-43 + the callers' F.relu)."""
    y = F.conv2d(x, conv.weight, None, conv.stride, conv.padding, conv.dilation)
    y = F.batch_norm(y, bn.running_mean, bn.running_var, bn.weight, bn.bias, bn.training, bn.momentum, bn.eps)
    if bn.training and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    return F.relu(y) if relu else y


def conv_bias(x, conv: nn.Conv2d):
    return F.conv2d(x, conv.weight, conv.bias, conv.stride, conv.padding, conv.dilation)


def concat(xs):
    return torch.cat(xs, 1)


def mul(a, b):
    return a * b


def add_relu(a, b):
    return F.relu(a + b)


def up2_align_corners(x):
    """nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True) (pranet.py:93)."""
    return ops.interpolate_bilinear(x, scale_factor=2, align_corners=True)
