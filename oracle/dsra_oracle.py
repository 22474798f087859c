"""CPU oracle for the DSRA decoder head and the dual-supervision losses.

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` leg may import this module, and only as
the checker (or the timed CPU baseline); the product package `pranet_v2_b200` never
imports anything from `oracle/`.

It is a *functional restatement* (state_dict in, tensors out; no nn.Module) of the
reference's algorithm for the hot path, fp32 on CPU, each function citing the reference
file:line it follows.  The reference's arithmetic lives in PyTorch ATen (pinned
torch==2.0.1, `pranet2.yaml:138`; installed 2.11.0 -- the semantics of conv2d, batch_norm,
bilinear interpolate at exact integer ratios, softmax, avg_pool2d and BCE-with-logits are
unchanged between the two), so the restatement calls the same ATen ops on CPU; the
element-wise pieces (bilinear sampling, boundary weight, structure loss) are additionally
restated in explicit float64 numpy (`*_np`) so the checker does not rest on ATen alone.

PARITY PIN: the reference ships no test, golden vector or fixture for this path
(SURVEY.md §4, §8c).  The pin is therefore created here: `oracle/make_golden.py` imports
the UNMODIFIED reference modules from /root/reference in the builder container, runs them
on the seeded inputs / weights of `oracle/synth.py` and freezes the outputs under
`tests/golden/`; `tests/test_oracle_golden.py` checks this oracle against those files.
"""
from __future__ import annotations

import itertools

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5       # nn.BatchNorm2d default, binary_seg/lib/pranet.py:37
BN_MOMENTUM = 0.1


# --------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------
def basic_conv(x, sd, prefix, training=False, padding=0, dilation=1):
    """BasicConv2d.forward: BN(conv(x)), conv bias=False, NO ReLU (binary_seg/lib/pranet.py:31-43).

    In training mode the running stats inside `sd` are updated in place exactly like
    nn.BatchNorm2d (momentum 0.1, unbiased variance) and num_batches_tracked is bumped."""
    y = F.conv2d(x, sd[prefix + ".conv.weight"], None, 1, padding, dilation)
    if training and (prefix + ".bn.num_batches_tracked") in sd:
        sd[prefix + ".bn.num_batches_tracked"] += 1
    return F.batch_norm(y, sd[prefix + ".bn.running_mean"], sd[prefix + ".bn.running_var"],
                        sd[prefix + ".bn.weight"], sd[prefix + ".bn.bias"], training, BN_MOMENTUM, BN_EPS)


def up2_ac(x):
    """nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True) (pranet.py:93)."""
    return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)


def interp(x, scale=None, size=None):
    """F.interpolate(..., mode='bilinear') with align_corners=None->False (pranet.py:349-415;
    EMCAD/lib/decoders.py:460-461 use the size= form)."""
    if size is not None:
        return F.interpolate(x, size=size, mode="bilinear")
    return F.interpolate(x, scale_factor=scale, mode="bilinear")


def rfb(x, sd, p, training=False):
    """RFB_modified.forward (pranet.py:46-83)."""
    bc = lambda t, name, **kw: basic_conv(t, sd, f"{p}.{name}", training, **kw)
    x0 = bc(x, "branch0.0")
    outs = [x0]
    for b, k in ((1, 3), (2, 5), (3, 7)):
        t = bc(x, f"branch{b}.0")
        t = bc(t, f"branch{b}.1", padding=(0, k // 2))
        t = bc(t, f"branch{b}.2", padding=(k // 2, 0))
        t = bc(t, f"branch{b}.3", padding=k, dilation=k)
        outs.append(t)
    x_cat = bc(torch.cat(outs, 1), "conv_cat", padding=1)
    return F.relu(x_cat + bc(x, "conv_res"))


def _aggregation_trunk(x1, x2, x3, sd, p, training):
    """Shared part of aggregation.forward (pranet.py:109-121 == PraNet_Res2Net.py:83-95)."""
    bc = lambda t, name: basic_conv(t, sd, f"{p}.{name}", training, padding=1)
    x1_1 = x1
    x2_1 = bc(up2_ac(x1), "conv_upsample1") * x2
    x3_1 = bc(up2_ac(up2_ac(x1)), "conv_upsample2") * bc(up2_ac(x2), "conv_upsample3") * x3
    x2_2 = torch.cat((x2_1, bc(up2_ac(x1_1), "conv_upsample4")), 1)
    x2_2 = bc(x2_2, "conv_concat2")
    x3_2 = torch.cat((x3_1, bc(up2_ac(x2_2), "conv_upsample5")), 1)
    x3_2 = bc(x3_2, "conv_concat3")
    return bc(x3_2, "conv4")


def aggregation_v2(x1, x2, x3, sd, p="agg1", training=False):
    """V2 aggregation: conv5_fg / conv5_bg are 1x1 WITH bias (pranet.py:103-104,122-123)."""
    x = _aggregation_trunk(x1, x2, x3, sd, p, training)
    return (F.conv2d(x, sd[p + ".conv5_fg.weight"], sd[p + ".conv5_fg.bias"]),
            F.conv2d(x, sd[p + ".conv5_bg.weight"], sd[p + ".conv5_bg.bias"]))


def aggregation_v1(x1, x2, x3, sd, p="agg1", training=False):
    """V1 aggregation: single conv5 1x1 with bias (PraNet_Res2Net.py:81,96)."""
    x = _aggregation_trunk(x1, x2, x3, sd, p, training)
    return F.conv2d(x, sd[p + ".conv5.weight"], sd[p + ".conv5.bias"])


def dsra_fuse(fg, crop_fg, crop_bg, use_softmax=True):
    """The V2 attention fusion (pranet.py:365-368): fg + fg * softmax_c(crop_fg - crop_bg)."""
    d = crop_fg - crop_bg
    if use_softmax:
        d = F.softmax(d, dim=1)
    return fg + fg.mul(d)


def ra_v1_scale(crop, x):
    """V1 reverse attention (PraNet_Res2Net.py:153-154): (1 - sigmoid(crop)).expand(C) * x."""
    a = -1 * torch.sigmoid(crop) + 1
    return a.expand(-1, x.shape[1], -1, -1).mul(x)


# --------------------------------------------------------------------------------------
# heads
# --------------------------------------------------------------------------------------
def _ra_stack(x, sd, stage, nconv, k, training):
    """ra{stage}_conv1 (no ReLU) then ra{stage}_conv2..n with ReLU (pranet.py:357-360,378-380,400-402)."""
    t = basic_conv(x, sd, f"ra{stage}_conv1", training)
    for i in range(2, nconv + 1):
        t = F.relu(basic_conv(t, sd, f"ra{stage}_conv{i}", training, padding=k // 2))
    return t


def pranet_v2_head(x2, x3, x4, sd, use_softmax=True, sem_downsample=1, training=False):
    """PraNet_V2.forward / PVT_PraNet_V2.forward after the backbone (pranet.py:343-417, 195-263).

    Returns the 8-tuple (l2_fg, l3_fg, l4_fg, l5_fg, l2_bg, l3_bg, l4_bg, l5_bg)."""
    x2_rfb = rfb(x2, sd, "rfb2_1", training)
    x3_rfb = rfb(x3, sd, "rfb3_1", training)
    x4_rfb = rfb(x4, sd, "rfb4_1", training)
    ra5_fg, ra5_bg = aggregation_v2(x4_rfb, x3_rfb, x2_rfb, sd, "agg1", training)
    l5_fg = interp(ra5_fg, 8 / sem_downsample)
    l5_bg = interp(ra5_bg, 8 / sem_downsample)
    # DSRA3
    crop_fg, crop_bg = interp(ra5_fg, 0.25), interp(ra5_bg, 0.25)
    t = _ra_stack(x4, sd, 4, 4, 5, training)
    fg = basic_conv(t, sd, "ra4_conv5_fg", training)
    bg = basic_conv(t, sd, "ra4_conv5_bg", training)
    fg = dsra_fuse(fg, crop_fg, crop_bg, use_softmax)
    l4_fg, l4_bg = interp(fg, 32 / sem_downsample), interp(bg, 32 / sem_downsample)
    # DSRA2
    crop_fg, crop_bg = interp(fg, 2), interp(bg, 2)
    t = _ra_stack(x3, sd, 3, 3, 3, training)
    fg = basic_conv(t, sd, "ra3_conv4_fg", training, padding=1)
    bg = basic_conv(t, sd, "ra3_conv4_bg", training, padding=1)
    fg = dsra_fuse(fg, crop_fg, crop_bg, use_softmax)
    l3_fg, l3_bg = interp(fg, 16 / sem_downsample), interp(bg, 16 / sem_downsample)
    # DSRA1
    crop_fg, crop_bg = interp(fg, 2), interp(bg, 2)
    t = _ra_stack(x2, sd, 2, 3, 3, training)
    fg = basic_conv(t, sd, "ra2_conv4_fg", training, padding=1)
    bg = basic_conv(t, sd, "ra2_conv4_bg", training, padding=1)
    fg = dsra_fuse(fg, crop_fg, crop_bg, use_softmax)
    l2_fg, l2_bg = interp(fg, 8 / sem_downsample), interp(bg, 8 / sem_downsample)
    return l2_fg, l3_fg, l4_fg, l5_fg, l2_bg, l3_bg, l4_bg, l5_bg


def pranet_v1_head(x2, x3, x4, sd, training=False):
    """V1 PraNet.forward after the backbone (PraNet_Res2Net.py:143-186) -> (l5, l4, l3, l2)."""
    x2_rfb = rfb(x2, sd, "rfb2_1", training)
    x3_rfb = rfb(x3, sd, "rfb3_1", training)
    x4_rfb = rfb(x4, sd, "rfb4_1", training)
    ra5 = aggregation_v1(x4_rfb, x3_rfb, x2_rfb, sd, "agg1", training)
    l5 = interp(ra5, 8)
    crop = interp(ra5, 0.25)
    t = _ra_stack(ra_v1_scale(crop, x4), sd, 4, 4, 5, training)
    x = basic_conv(t, sd, "ra4_conv5", training) + crop
    l4 = interp(x, 32)
    crop = interp(x, 2)
    t = _ra_stack(ra_v1_scale(crop, x3), sd, 3, 3, 3, training)
    x = basic_conv(t, sd, "ra3_conv4", training, padding=1) + crop
    l3 = interp(x, 16)
    crop = interp(x, 2)
    t = _ra_stack(ra_v1_scale(crop, x2), sd, 2, 3, 3, training)
    x = basic_conv(t, sd, "ra2_conv4", training, padding=1) + crop
    l2 = interp(x, 8)
    return l5, l4, l3, l2


def dual_heads_cascade(feats, sd, kernel_sizes=(1, 3, 3, 3), use_softmax=True, training=False,
                       names=("ConvBlock4", "ConvBlock3", "ConvBlock2", "ConvBlock1"), bn=True):
    """The DSRA part of the multiclass host decoders, given the decoder features d4..d1
    (deep -> shallow): per stage fg/bg heads on the same feature, deeper fg/bg resized with
    F.interpolate(size=), softmax fusion of the fg map.
    EMCAD_dual.forward (EMCAD/lib/decoders.py:454-526), CASCADE_Add_dual.forward
    (MERIT/lib/decoders.py:342-431); with bn=False the heads are plain 1x1 convs with bias as in
    MIST CAM (MIST/lib/MIST.py:403-451, names out_head{1..4})."""
    fgs, bgs = [], []
    for i, (d, k, n) in enumerate(zip(feats, kernel_sizes, names)):
        if bn:
            fg = basic_conv(d, sd, n + "_fg", training, padding=k // 2)
            bg = basic_conv(d, sd, n + "_bg", training, padding=k // 2)
        else:
            fg = F.conv2d(d, sd[n + "_fg.weight"], sd[n + "_fg.bias"])
            bg = F.conv2d(d, sd[n + "_bg.weight"], sd[n + "_bg.bias"])
        if i > 0:
            up_fg = interp(fgs[-1], size=d.shape[2:])
            up_bg = interp(bgs[-1], size=d.shape[2:])
            fg = dsra_fuse(fg, up_fg, up_bg, use_softmax)
        fgs.append(fg)
        bgs.append(bg)
    return fgs + bgs


def final_upsample(maps, scales=(32, 16, 8, 4)):
    """EMCADNet.forward dual branch (EMCAD/lib/networks.py:114-125)."""
    n = len(scales)
    return [interp(m, scales[i % n]) for i, m in enumerate(maps)]


# --------------------------------------------------------------------------------------
# inference tails
# --------------------------------------------------------------------------------------
def infer_tail_binary(lowres_maps, scale_factors, size=None):
    """Test-time post-processing of binary_seg/MyTest_med.py:35-42 (:104-111; V1 :98-102 with one map), starting from the
    low-res foreground maps: final upsamples of the model (pranet.py:349-415), p2+p3+p4+p5, resize to the ground-truth
    size, sigmoid, min-max normalisation, uint8.  Per image (the reference runs batch 1).  Returns uint8 (B, H, W)."""
    ups = [interp(m, s) for m, s in zip(lowres_maps, scale_factors)]
    out = ups[0]
    for u in ups[1:]:
        out = out + u
    if size is None:
        size = tuple(out.shape[-2:])
    res = []
    for b in range(out.shape[0]):
        o = F.interpolate(out[b:b + 1], size=tuple(size), mode="bilinear", align_corners=False)
        o = o.sigmoid().data.cpu().numpy().squeeze()
        o = (o - o.min()) / (o.max() - o.min() + 1e-8)
        res.append((o * 255).astype(np.uint8))
    return np.stack(res)


def infer_tail_binary_float(lowres_maps, scale_factors, size=None):
    """Same as infer_tail_binary but returning the float value before the uint8 truncation (to tell +-1 truncation flips
    from real errors)."""
    ups = [interp(m, s) for m, s in zip(lowres_maps, scale_factors)]
    out = ups[0]
    for u in ups[1:]:
        out = out + u
    if size is None:
        size = tuple(out.shape[-2:])
    res = []
    for b in range(out.shape[0]):
        o = F.interpolate(out[b:b + 1], size=tuple(size), mode="bilinear", align_corners=False)
        o = o.sigmoid().data.cpu().numpy().squeeze()
        res.append((o - o.min()) / (o.max() - o.min() + 1e-8) * 255)
    return np.stack(res)


def infer_tail_argmax(P_fg_low, P_bg_low, scales=(32, 16, 8, 4)):
    """Dual-branch prediction rule of EMCAD/utils/utils.py:261-273,285-296 from the low-res stage maps: final upsamples of
    EMCAD/lib/networks.py:116-123, outputs = sum_k (P[k] - P_bg[k]), argmax_c softmax.  Returns (labels uint8 (B,H,W),
    margin (B,H,W) = top-1 minus top-2 of the summed logits, so near-ties can be excluded from exact comparisons)."""
    outputs = 0.0
    for fg, bg, s in zip(P_fg_low, P_bg_low, scales):
        outputs += (interp(fg, s) - interp(bg, s))
    out = torch.argmax(torch.softmax(outputs, dim=1), dim=1)
    top2 = torch.topk(outputs, 2, dim=1).values if outputs.shape[1] > 1 else torch.cat([outputs, outputs - 1], 1)
    return out.numpy().astype(np.uint8), (top2[:, 0] - top2[:, 1]).numpy()


# --------------------------------------------------------------------------------------
# losses
# --------------------------------------------------------------------------------------
def structure_loss(pred, pred_bg, mask_fg, mask_bg):
    """binary_seg/MyTrain_med.py:19-38."""
    weit = 1 + 5 * torch.abs(F.avg_pool2d(mask_fg, kernel_size=31, stride=1, padding=15) - mask_fg)
    wbce = F.binary_cross_entropy_with_logits(pred, mask_fg, reduction="none")
    wbce = (weit * wbce).sum(dim=(2, 3)) / weit.sum(dim=(2, 3))
    wbce2 = F.binary_cross_entropy_with_logits(pred_bg, mask_bg, reduction="none")
    wbce2 = (weit * wbce2).sum(dim=(2, 3)) / weit.sum(dim=(2, 3))
    p = torch.sigmoid(pred)
    inter = ((p * mask_fg) * weit).sum(dim=(2, 3))
    union = ((p + mask_fg) * weit).sum(dim=(2, 3))
    wiou = 1 - (inter + 1) / (union - inter + 1)
    return (wbce + wiou + 0.8 * wbce2).mean()


def structure_loss_lowres(pairs, scale_factors, mask_fg, mask_bg=None):
    """The final upsamples of PraNet_V2.forward (binary_seg/lib/pranet.py:349-350,370-371,392-393,414-415) followed by the loss calls
    of the training loop (MyTrain_med.py:74,78-81), starting from the low-res (fg, bg) pairs: tensor of len(pairs) losses, in pair order
    (SURVEY.md 8 f2 -- what pv2_structure_loss_lowres_fwd computes in one pass)."""
    if mask_bg is None:
        mask_bg = 1 - mask_fg                                                    # MyTrain_med.py:74
    return torch.stack([structure_loss(interp(a, scale=s), interp(b, scale=s), mask_fg, mask_bg) for (a, b), s in zip(pairs, scale_factors)])


def powerset_subsets(n=4):
    """Non-empty subsets in the order the reference's powerset() generator yields them
    (EMCAD/utils/utils.py:20-30; the empty set is skipped at trainer.py:132-133)."""
    def gen(seq):
        if len(seq) <= 1:
            yield seq
            yield []
        else:
            for item in gen(seq[1:]):
                yield [seq[0]] + item
                yield item
    return [s for s in gen(list(range(n))) if s]


def inverted_one_hot(labels, num_classes):
    """convert_labels_to_one_hot_masks (EMCAD/trainer.py:22-29): 1 - onehot(label), float."""
    oh = F.one_hot(labels.long(), num_classes).permute(0, 3, 1, 2)
    return (1 - oh).float()


def dice_loss(logits, labels, num_classes):
    """DiceLoss.forward(softmax=True) (EMCAD/utils/utils.py:102-138): batch-global per-class dice."""
    p = torch.softmax(logits, dim=1)
    loss = 0.0
    for c in range(num_classes):
        t = (labels == c).float()
        s = p[:, c]
        inter = torch.sum(s * t)
        loss = loss + (1 - (2 * inter + 1e-5) / (torch.sum(s * s) + torch.sum(t * t) + 1e-5))
    return loss / num_classes


def mc_dual_loss(P_fg, P_bg, labels, num_classes, subsets=None, lc=(0.5, 0.7, 0.3)):
    """The multiclass dual-supervision loss (EMCAD/trainer.py:123-140 == MERIT/train_ACDC.py:259-284
    == MIST/trainer.py:112-129): sum over the 15 non-empty subsets s of the 4 scales of
    0.5*CE(sum fg[s], y) + 0.7*Dice(softmax(sum fg[s]), y) + 0.3*BCEWithLogits(sum bg[s], 1-onehot(y))."""
    subsets = subsets if subsets is not None else powerset_subsets(len(P_fg))
    bg_mask = inverted_one_hot(labels, num_classes)
    loss = 0.0
    for s in subsets:
        iout = sum(P_fg[i] for i in s)
        ibg = sum(P_bg[i] for i in s)
        loss = loss + lc[0] * F.cross_entropy(iout, labels.long()) \
            + lc[1] * dice_loss(iout, labels, num_classes) \
            + lc[2] * F.binary_cross_entropy_with_logits(ibg, bg_mask)
    return loss


# --------------------------------------------------------------------------------------
# explicit float64 numpy restatements of the element-wise pieces
# --------------------------------------------------------------------------------------
def _src_index(out_size, in_size, align_corners, scale_factor=None):
    """ATen area_pixel_compute_source_index for bilinear (cubic=False).
    align_corners=False: src = (dst+0.5)*r - 0.5 clamped at 0, r = 1/scale_factor if one was
    given (recompute_scale_factor=None keeps the user's factor) else in/out."""
    d = np.arange(out_size, dtype=np.float64)
    if align_corners:
        r = (in_size - 1) / (out_size - 1) if out_size > 1 else 0.0
        src = d * r
    else:
        r = (1.0 / scale_factor) if scale_factor else in_size / out_size
        src = np.maximum((d + 0.5) * r - 0.5, 0.0)
    i0 = np.minimum(np.floor(src).astype(np.int64), in_size - 1)
    i1 = i0 + (i0 < in_size - 1)
    l1 = src - i0
    return i0, i1, 1.0 - l1, l1


def bilinear_np(x, out_h, out_w, align_corners=False, scale_factor=None):
    x = np.asarray(x, np.float64)
    y0, y1, wy0, wy1 = _src_index(out_h, x.shape[2], align_corners, scale_factor)
    x0, x1, wx0, wx1 = _src_index(out_w, x.shape[3], align_corners, scale_factor)
    top = x[:, :, y0][:, :, :, x0] * wx0 + x[:, :, y0][:, :, :, x1] * wx1
    bot = x[:, :, y1][:, :, :, x0] * wx0 + x[:, :, y1][:, :, :, x1] * wx1
    return top * wy0[None, None, :, None] + bot * wy1[None, None, :, None]


def boundary_weight_np(mask):
    """weit = 1 + 5*|avgpool31(mask) - mask|, zero padding COUNTED in the divisor (961)
    (MyTrain_med.py:21; avg_pool2d default count_include_pad=True)."""
    m = np.asarray(mask, np.float64)
    p = np.pad(m, ((0, 0), (0, 0), (15, 15), (15, 15)))
    c = np.cumsum(np.cumsum(p, axis=2), axis=3)
    c = np.pad(c, ((0, 0), (0, 0), (1, 0), (1, 0)))
    H, W = m.shape[2], m.shape[3]
    box = c[:, :, 31:31 + H, 31:31 + W] - c[:, :, 0:H, 31:31 + W] - c[:, :, 31:31 + H, 0:W] + c[:, :, 0:H, 0:W]
    return 1.0 + 5.0 * np.abs(box / 961.0 - m)


def structure_loss_np(pred, pred_bg, mask_fg, mask_bg=None):
    """float64 restatement of MyTrain_med.py:19-38; returns (loss, dpred, dpred_bg)."""
    x = np.asarray(pred, np.float64)
    xb = np.asarray(pred_bg, np.float64)
    m = np.asarray(mask_fg, np.float64)
    mb = 1.0 - m if mask_bg is None else np.asarray(mask_bg, np.float64)
    w = boundary_weight_np(m)
    sp = lambda z: np.maximum(z, 0) + np.log1p(np.exp(-np.abs(z)))  # softplus
    bce = sp(x) - x * m
    bce2 = sp(xb) - xb * mb
    s = 1.0 / (1.0 + np.exp(-x))
    ax = (2, 3)
    W = w.sum(ax)
    inter = (s * m * w).sum(ax)
    union = ((s + m) * w).sum(ax)
    den = union - inter + 1
    per = (w * bce).sum(ax) / W + 1 - (inter + 1) / den + 0.8 * (w * bce2).sum(ax) / W
    n = per.size
    loss = per.mean()
    # gradients
    Wb, ib, db = W[..., None, None], inter[..., None, None], den[..., None, None]
    ds = s * (1 - s)
    # d wiou / d s_ij = -[ m w den - (inter+1) (w - m w) ] / den^2
    dwiou = -((m * w) * db - (ib + 1) * (w - m * w)) / (db * db)
    dpred = (w * (s - m) / Wb + dwiou * ds) / n
    sb = 1.0 / (1.0 + np.exp(-xb))
    dpred_bg = 0.8 * w * (sb - mb) / Wb / n
    return loss, dpred, dpred_bg
