"""Case table shared by `oracle/make_golden.py` (writes tests/golden/*.npz from the
REFERENCE) and the tests (re-generate the same seeded inputs, compare).  TEST INFRASTRUCTURE.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# ---- structure_loss (binary_seg/MyTrain_med.py:19-38) --------------------------------
# name -> (B, C, H, W, mask kind, sample stride used when storing gradients)
STRUCTURE_LOSS_CASES = {
    "sl_hard_64": (2, 1, 64, 64, "hard", 1),
    "sl_soft_96x80": (2, 1, 96, 80, "soft", 1),
    "sl_c2_40x56": (3, 2, 40, 56, "soft", 1),
    "sl_tiny_16x20": (1, 1, 16, 20, "hard", 1),      # smaller than the 31x31 window
    "sl_31": (1, 1, 31, 31, "hard", 1),
    "sl_empty_mask": (2, 1, 48, 48, "zeros", 1),
    "sl_full_mask": (1, 1, 48, 48, "ones", 1),
    "sl_hard_352": (2, 1, 352, 352, "hard", 11),
    "sl_soft_448": (1, 1, 448, 448, "soft", 13),
}


# loss from the LOW-RES head maps: (B, S, sem_downsample, mask kind); the maps are the eight tensors PraNet_V2.forward
# holds just before its final F.interpolate calls (pranet.py:349-350,370-371,392-393,414-415): S/8, S/16, S/32, S/8 (over sem_downsample)
LOWRES_LOSS_CASES = {
    "ll_352": (2, 352, 1, "hard"),
    "ll_soft_256": (2, 256, 1, "soft"),          # multi-scale rate 0.75: resized (soft) masks, MyTrain_med.py:70-73
    "ll_sd2_128": (1, 128, 2, "hard"),           # sem_downsample=2: final factors 4, 8, 16, 4
    "ll_160": (3, 160, 1, "hard"),               # W % 128 != 0: partial tiles
}


def lowres_loss_scales(name):
    sd = LOWRES_LOSS_CASES[name][2]
    return [8 / sd, 16 / sd, 32 / sd, 8 / sd]


def lowres_loss_inputs(name):
    """-> (maps: [fg2, fg3, fg4, ra5_fg, bg2, bg3, bg4, ra5_bg] in the model's output order, mask (B,1,S/sd,S/sd))."""
    B, S, sd, kind = LOWRES_LOSS_CASES[name]
    seed = abs(hash_name(name))
    maps = [synth.logits((B, 1, S // d, S // d), seed, f"low{i}") for i, d in enumerate((8, 16, 32, 8, 8, 16, 32, 8))]
    H = S // sd
    m = (synth.ellipse_masks(B, H, H, seed) if kind == "hard" else synth.soft_masks(B, H, H, seed)).view(B, 1, H, H)
    return maps, m.contiguous()


def structure_loss_inputs(name):
    B, C, H, W, kind, _ = STRUCTURE_LOSS_CASES[name]
    seed = abs(hash_name(name))
    pred = synth.logits((B, C, H, W), seed, "pred")
    pred_bg = synth.logits((B, C, H, W), seed, "pred_bg")
    if kind == "hard":
        m = synth.ellipse_masks(B * C, H, W, seed).view(B, C, H, W)
    elif kind == "soft":
        m = synth.soft_masks(B * C, H, W, seed, base=max(8, (min(H, W) * 4 // 5) // 8 * 8)).view(B, C, H, W)
    elif kind == "zeros":
        m = torch.zeros(B, C, H, W)
    else:
        m = torch.ones(B, C, H, W)
    return pred, pred_bg, m.contiguous(), (1 - m).contiguous()


def hash_name(name: str) -> int:
    import zlib
    return zlib.crc32(name.encode()) % 100000


# ---- PraNet heads on synthetic backbone features ------------------------------------
# name -> dict(model, kwargs, channels, B, size, training, out_stride)
HEAD_CASES = {
    "v2_res_c1_eval": dict(model="PraNet_V2", kw=dict(num_class=1), ch=synth.RES2NET_CH, B=2, size=64, training=False, stride=1),
    "v2_res_c1_train": dict(model="PraNet_V2", kw=dict(num_class=1), ch=synth.RES2NET_CH, B=2, size=64, training=True, stride=1),
    "v2_res_c3_train": dict(model="PraNet_V2", kw=dict(num_class=3), ch=synth.RES2NET_CH, B=2, size=96, training=True, stride=3),
    "v2_res_c3_eval": dict(model="PraNet_V2", kw=dict(num_class=3), ch=synth.RES2NET_CH, B=1, size=96, training=False, stride=3),
    "v2_res_c3_linear": dict(model="PraNet_V2", kw=dict(num_class=3, use_softmax=False), ch=synth.RES2NET_CH, B=1, size=64, training=False, stride=2),
    "v2_res_c1_sem2": dict(model="PraNet_V2", kw=dict(num_class=1, sem_downsample=2), ch=synth.RES2NET_CH, B=1, size=64, training=False, stride=1),
    "v2_pvt_c1_eval": dict(model="PVT_PraNet_V2", kw=dict(num_class=1), ch=synth.PVT_CH, B=2, size=64, training=False, stride=1),
    "v2_pvt_c1_train": dict(model="PVT_PraNet_V2", kw=dict(num_class=1), ch=synth.PVT_CH, B=2, size=96, training=True, stride=2),
    "v1_res_eval": dict(model="PraNet", kw=dict(), ch=synth.RES2NET_CH, B=2, size=64, training=False, stride=1),
    "v1_res_train": dict(model="PraNet", kw=dict(), ch=synth.RES2NET_CH, B=2, size=96, training=True, stride=2),
    "v1_pvt_eval": dict(model="PVT_PraNet", kw=dict(), ch=synth.PVT_CH, B=1, size=64, training=False, stride=1),
}
# cases for which the training-step gradients (4x structure_loss, MyTrain_med.py:78-82) are pinned too
HEAD_GRAD_CASES = ("v2_res_c1_train", "v2_pvt_c1_train")


def head_inputs(name):
    c = HEAD_CASES[name]
    seed = hash_name(name)
    return synth.backbone_features(c["B"], c["size"], seed, c["ch"])


def head_key_filter(k: str) -> bool:
    """state_dict keys that belong to the head (everything but the backbone / grayscale stem)."""
    return not (k.startswith("backbone.") or k.startswith("resnet.") or k.startswith("conv."))


# ---- multiclass DSRA carriers ---------------------------------------------------------
MC_CASES = {
    "emcad_c9": dict(kind="emcad", channels=[512, 320, 128, 64], num_class=9, B=2, size=64, training=True),
    "emcad_c4_eval": dict(kind="emcad", channels=[512, 320, 128, 64], num_class=4, B=1, size=96, training=False),
    "merit_c4": dict(kind="merit", channels=[768, 384, 192, 96], num_class=4, B=2, size=64, training=True),
    "merit_c9_linear": dict(kind="merit", channels=[768, 384, 192, 96], num_class=9, B=1, size=64, training=False, use_softmax=False),
    "mist_c9": dict(kind="mist", channels=[768, 384, 192, 96], num_class=9, B=1, size=64, training=False),
}

# ---- the drop-in decoder classes (pranet_v2_b200.multiclass) against the reference decoders, every weight synthetic ----
MC_DEC_CASES = {
    "mcdec_emcad_c9": dict(kind="emcad", channels=[512, 320, 128, 64], num_class=9, B=2, size=64, kw=dict(expansion_factor=2, activation="relu")),
    "mcdec_merit_c4": dict(kind="merit", channels=[768, 384, 192, 96], num_class=4, B=2, size=64, kw=dict(use_softmax=True)),
}


def mc_dec_pyramid(name):
    c = MC_DEC_CASES[name]
    seed = hash_name(name)
    return [torch.randn(c["B"], ch, c["size"] // s, c["size"] // s, generator=synth._gen(seed, f"pyr{i}"))
            for i, (ch, s) in enumerate(zip(c["channels"], (32, 16, 8, 4)))]


def mc_dec_out_weights(name, shapes):
    """Fixed random cotangents: the scalar that is backpropagated is sum_i <w_i, out_i>."""
    seed = hash_name(name)
    return [torch.randn(tuple(sh), generator=synth._gen(seed, f"cot{i}")) for i, sh in enumerate(shapes)]


# ---- multiclass dual loss -------------------------------------------------------------
MC_LOSS_CASES = {
    "mcl_c9_32": dict(num_class=9, B=2, H=32, W=32),
    "mcl_c4_24x40": dict(num_class=4, B=3, H=24, W=40),
}


def mc_loss_inputs(name):
    c = MC_LOSS_CASES[name]
    seed = hash_name(name)
    shape = (c["B"], c["num_class"], c["H"], c["W"])
    P_fg = [synth.logits(shape, seed, f"fg{i}", 1.5) for i in range(4)]
    P_bg = [synth.logits(shape, seed, f"bg{i}", 1.5) for i in range(4)]
    labels = synth.class_labels(c["B"], c["H"], c["W"], c["num_class"], seed)
    return P_fg, P_bg, labels


# ---- inference tails from the low-res maps (binary_seg/MyTest_med.py:35-42; EMCAD/utils/utils.py:285-296) ----
# binary: input size S (maps at S/8, S/16, S/32, S/8), ground-truth size the reference resizes to; nmaps=1 is the V1 rule
TAIL_BINARY_CASES = {
    "tail_bin_352_same": dict(B=2, S=352, gt=(352, 352), nmaps=4),
    "tail_bin_352_to_288x384": dict(B=1, S=352, gt=(288, 384), nmaps=4),
    "tail_bin_64_to_531x473": dict(B=2, S=64, gt=(531, 473), nmaps=4),     # odd, non-multiple-of-4 ground-truth size
    "tail_bin_v1_96_to_100x120": dict(B=1, S=96, gt=(100, 120), nmaps=1),
}
TAIL_MC_CASES = {
    "tail_mc_c9_224": dict(B=2, C=9, S=224),
    "tail_mc_c4_96": dict(B=1, C=4, S=96),
}
TAIL_BIN_SCALES = (8, 16, 32, 8)
TAIL_MC_SCALES = (32, 16, 8, 4)


def tail_binary_inputs(name):
    c = TAIL_BINARY_CASES[name]
    seed = hash_name(name)
    scales = TAIL_BIN_SCALES[:c["nmaps"]]
    return [synth.logits((c["B"], 1, c["S"] // s, c["S"] // s), seed, f"m{k}", 2.0) for k, s in enumerate(scales)], list(scales)


def tail_mc_inputs(name):
    c = TAIL_MC_CASES[name]
    seed = hash_name(name)
    fg = [synth.logits((c["B"], c["C"], c["S"] // s, c["S"] // s), seed, f"fg{k}", 2.0) for k, s in enumerate(TAIL_MC_SCALES)]
    bg = [synth.logits((c["B"], c["C"], c["S"] // s, c["S"] // s), seed, f"bg{k}", 2.0) for k, s in enumerate(TAIL_MC_SCALES)]
    return fg, bg


# ---- full models (stock backbone + head) ---------------------------------------------
FULL_CASES = {
    "full_v2_res_352": dict(model="PraNet_V2", kw=dict(num_class=1), B=1, size=352, training=True, stride=4),
    "full_v2_res_64_eval": dict(model="PraNet_V2", kw=dict(num_class=1), B=2, size=64, training=False, stride=1),
    "full_v2_pvt_64": dict(model="PVT_PraNet_V2", kw=dict(num_class=1), B=2, size=64, training=False, stride=1),   # eval: DropPath is stochastic in train
    "full_v1_res_128": dict(model="PraNet", kw=dict(), B=2, size=128, training=False, stride=2),   # eval: tiny-batch train-mode BN on 4x4 maps amplifies backbone rounding
}


def full_input(name):
    c = FULL_CASES[name]
    g = torch.Generator().manual_seed(hash_name(name))
    return torch.randn(c["B"], 3, c["size"], c["size"], generator=g)


def load(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        return {k: z[k] for k in z.files}
