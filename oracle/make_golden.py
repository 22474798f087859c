"""Freeze golden vectors from the UNMODIFIED reference (builder container only).

    python -m oracle.make_golden [group ...]      # groups: tail sl ll head mc mcl full  (default: all)

Imports the reference's own modules from /root/reference through `oracle/ref_import.py`,
feeds them the seeded inputs / weights of `oracle/synth.py` + `oracle/golden_cases.py`, and
writes the OUTPUTS to tests/golden/<case>.npz.  The reference cannot travel to the GPU box,
these files do.  TEST INFRASTRUCTURE -- never imported by the product.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import golden_cases as G
from . import ref_import as R
from . import synth

torch.set_num_threads(max(1, os.cpu_count() or 1))


def _save(name, **arrays):
    os.makedirs(G.GOLDEN_DIR, exist_ok=True)
    path = os.path.join(G.GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrays.items()})
    print(f"  wrote {name}.npz ({os.path.getsize(path) / 1024:.0f} KiB)")


def _np(t, stride=1):
    t = t.detach().cpu()
    if stride > 1 and t.ndim == 4:
        t = t[:, :, ::stride, ::stride]
    return t.contiguous().numpy()


# --------------------------------------------------------------------------------------
def gen_structure_loss():
    sl = R.structure_loss()
    for name, (B, C, H, W, kind, stride) in G.STRUCTURE_LOSS_CASES.items():
        pred, pred_bg, m, mb = G.structure_loss_inputs(name)
        pred.requires_grad_(True)
        pred_bg.requires_grad_(True)
        loss = sl(pred, pred_bg, m, mb)
        loss.backward()
        _save(name, loss=np.float64(loss.item()), dpred=_np(pred.grad, stride), dpred_bg=_np(pred_bg.grad, stride),
              mask_sum=np.float64(m.double().sum().item()))


def gen_lowres_loss():
    """Final upsamples as PraNet_V2.forward issues them (pranet.py:349-350,370-371,392-393,414-415: F.interpolate(scale_factor=,
    mode='bilinear')) followed by the reference's own loss statements (MyTrain_med.py:74,78-82), differentiated back to the low-res maps."""
    block = R.binary_train_loss_block()
    for name in G.LOWRES_LOSS_CASES:
        maps, m = G.lowres_loss_inputs(name)
        maps = [t.requires_grad_(True) for t in maps]
        ups = [F.interpolate(t, scale_factor=s, mode="bilinear") for t, s in zip(maps, G.lowres_loss_scales(name) * 2)]
        loss, (l2, l3, l4, l5) = block(tuple(ups), m)
        loss.backward()
        # MyTrain_med.py:78-81 pairs loss5 with lateral_map_2, loss4 with _3, loss3 with _4, loss2 with _5: store in MAP order
        _save(name, loss=np.float64(loss.item()), losses=np.array([l5.item(), l4.item(), l3.item(), l2.item()], np.float64),
              **{f"d{i}": _np(t.grad) for i, t in enumerate(maps)})


# --------------------------------------------------------------------------------------
class _StubRes2Net(nn.Module):
    """Stands in for the Res2Net backbone so the reference forward runs its head on given features
    (binary_seg/lib/pranet.py:331-341 calls conv1/bn1/relu/maxpool/layer1..4 one by one)."""

    def __init__(self, feats):
        super().__init__()
        self.conv1 = self.bn1 = self.relu = self.maxpool = self.layer1 = nn.Identity()
        self._f = feats

    def layer2(self, _):
        return self._f[0]

    def layer3(self, _):
        return self._f[1]

    def layer4(self, _):
        return self._f[2]


class _StubPVT(nn.Module):
    def __init__(self, feats):
        super().__init__()
        self._f = feats

    def forward(self, _):
        return None, self._f[0], self._f[1], self._f[2]


def _build_head_model(case, feats):
    m = R.build_binary(case["model"], **case["kw"])
    if hasattr(m, "resnet"):
        m.resnet = _StubRes2Net(feats)
    elif case["model"] in ("PVT_PraNet_V2", "PVT_PraNet"):
        m.backbone = _StubPVT(feats)
    else:
        m.backbone = _StubRes2Net(feats)
    tmpl = {k: v for k, v in m.state_dict().items() if G.head_key_filter(k)}
    sd = synth.synth_state_dict(tmpl, seed=1)
    missing = m.load_state_dict(sd, strict=False)
    assert not [k for k in missing.missing_keys if G.head_key_filter(k)], missing
    return m, sd


def gen_heads():
    sl = R.structure_loss()
    for name, case in G.HEAD_CASES.items():
        feats = G.head_inputs(name)
        want_grad = name in G.HEAD_GRAD_CASES
        if want_grad:
            feats = [f.requires_grad_(True) for f in feats]
        m, _ = _build_head_model(case, feats)
        m.train(case["training"])
        dummy = torch.zeros(case["B"], 3, 8, 8)
        outs = m(dummy)
        arrays = {f"out{i}": _np(o, case["stride"]) for i, o in enumerate(outs)}
        arrays["out_shape"] = np.array(outs[0].shape)
        if case["training"]:
            post = m.state_dict()
            for k in ("ra4_conv2.bn.running_mean", "ra4_conv2.bn.running_var", "rfb2_1.branch1.3.bn.running_var",
                      "agg1.conv4.bn.running_mean", "ra2_conv1.bn.running_var", "ra2_conv1.bn.num_batches_tracked"):
                arrays["stat:" + k] = _np(post[k])
        if want_grad:
            S = outs[0].shape[-1]
            gt = synth.ellipse_masks(case["B"], S, S, seed=7)
            loss = sum(sl(outs[i], outs[i + 4], gt, 1 - gt) for i in range(4))   # MyTrain_med.py:78-82
            loss.backward()
            arrays["loss"] = np.float64(loss.item())
            for i, f in enumerate(feats):
                arrays[f"dfeat{i}"] = _np(f.grad)
            for k, p in m.named_parameters():
                if G.head_key_filter(k) and p.grad is not None:
                    g = p.grad.double()
                    arrays["dw:" + k] = np.array([g.sum().item(), g.norm().item()] + g.flatten()[:6].tolist())
        _save(name, **arrays)


# --------------------------------------------------------------------------------------
def _pyramid(channels, B, size, seed):
    feats = []
    for i, (c, s) in enumerate(zip(channels, (32, 16, 8, 4))):
        feats.append(torch.randn(B, c, size // s, size // s, generator=synth._gen(seed, f"pyr{i}")))
    return feats


def gen_multiclass():
    for name, case in G.MC_CASES.items():
        seed = G.hash_name(name)
        ch, nc = case["channels"], case["num_class"]
        pyr = _pyramid(ch, case["B"], case["size"], seed)
        torch.manual_seed(seed)
        if case["kind"] == "emcad":
            dec = R.emcad_decoders().EMCAD_dual(channels=ch, num_class=nc)
            taps = [dec.mscb4, dec.mscb3, dec.mscb2, dec.mscb1]
            run = lambda: dec(pyr[0], pyr[1:])
        elif case["kind"] == "merit":
            dec = R.merit_decoders().CASCADE_Add_dual(channels=ch, num_class=nc, use_softmax=case.get("use_softmax", True))
            taps = [dec.ConvBlock4, dec.ConvBlock3, dec.ConvBlock2, dec.ConvBlock1]
            run = lambda: dec(pyr[0], pyr[1:])[:8]
        else:
            dec = R.mist_cam().CAM("SSS", channels=ch, n_class=nc)
            taps = [dec.block_6, dec.block_7, dec.block_8, dec.block_9]
            run = lambda: dec(pyr[3], pyr[2], pyr[1], pyr[0])
        # seeded weights for the DSRA heads only (the context blocks keep their own random init)
        head_keys = {k: v for k, v in dec.state_dict().items() if ("_fg." in k or "_bg." in k)}
        dec.load_state_dict(synth.synth_state_dict(head_keys, seed=2), strict=False)
        dec.train(case["training"])
        feats = []
        hooks = [t.register_forward_hook(lambda _m, _i, o: feats.append(o.detach().clone())) for t in taps]
        with torch.no_grad():
            outs = run()
        for h in hooks:
            h.remove()
        assert len(feats) == 4, len(feats)
        arrays = {f"d{i}": _np(f) for i, f in enumerate(feats)}
        arrays.update({f"out{i}": _np(o) for i, o in enumerate(outs)})
        # EMCADNet.forward dual branch: x32/x16/x8/x4 final upsample (EMCAD/lib/networks.py:114-125)
        import torch.nn.functional as F
        for i, o in enumerate(outs):
            arrays[f"up{i}"] = _np(F.interpolate(o, scale_factor=(32, 16, 8, 4)[i % 4], mode="bilinear"), 2)
        _save(name, **arrays)


def gen_mc_decoders():
    """The reference's dual decoders (EMCAD_dual, CASCADE_Add_dual) with EVERY weight synthetic (keyed by parameter name), train mode,
    forward + backward of a fixed linear functional of the outputs: what pranet_v2_b200.multiclass must reproduce from the same
    state_dict.  Also freezes the state_dict key / shape lists of the three decoder classes and of EMCADNet(dual=True)."""
    import json
    keys = {}
    for name, case in G.MC_DEC_CASES.items():
        ch, nc = case["channels"], case["num_class"]
        if case["kind"] == "emcad":
            dec = R.emcad_decoders().EMCAD_dual(channels=ch, num_class=nc, **case["kw"])
        else:
            dec = R.merit_decoders().CASCADE_Add_dual(channels=ch, num_class=nc, **case["kw"])
        dec.load_state_dict(synth.synth_state_dict(dec.state_dict(), seed=4))
        dec.train()
        pyr = [p.clone().requires_grad_(True) for p in G.mc_dec_pyramid(name)]
        outs = list(dec(pyr[0], pyr[1:]))[:8]
        cots = G.mc_dec_out_weights(name, [o.shape for o in outs])
        sum((o * w).sum() for o, w in zip(outs, cots)).backward()
        arrays = {f"out{i}": _np(o) for i, o in enumerate(outs)}
        arrays.update({f"dpyr{i}": _np(p.grad) for i, p in enumerate(pyr)})
        for k, p in dec.named_parameters():
            g = p.grad.double() if p.grad is not None else torch.zeros(1, dtype=torch.double)
            arrays["dw:" + k] = np.array([g.sum().item(), g.norm().item()])
        post = dec.state_dict()
        for k in post:
            if k.endswith(("running_mean", "running_var")) and ("_fg." in k or "_bg." in k):
                arrays["stat:" + k] = _np(post[k])
        _save(name, **arrays)
    # key / shape fixtures
    mk = lambda m: {k: list(v.shape) for k, v in m.state_dict().items()}
    keys["EMCAD_dual"] = mk(R.emcad_decoders().EMCAD_dual(channels=[512, 320, 128, 64], num_class=9))
    keys["CASCADE_Add_dual"] = mk(R.merit_decoders().CASCADE_Add_dual(channels=[768, 384, 192, 96], num_class=4))
    keys["CAM"] = mk(R.mist_cam().CAM("SSS", channels=[768, 384, 192, 96], n_class=9))
    try:
        net = R.emcad_networks().EMCADNet(num_classes=9, encoder="pvt_v2_b2", pretrain=False, dual=True)
        keys["EMCADNet"] = mk(net)
    except Exception as exc:      # noqa: BLE001
        print("  EMCADNet keys not frozen:", exc)
    path = os.path.join(G.GOLDEN_DIR, "mc_state_dict_keys.json")
    with open(path, "w") as f:
        json.dump(keys, f, sort_keys=True)
    print(f"  wrote mc_state_dict_keys.json ({os.path.getsize(path) / 1024:.0f} KiB, {sum(len(v) for v in keys.values())} keys)")


def gen_mc_loss():
    powerset, DiceLoss, to_inv = R.emcad_loss_pieces()
    for name, case in G.MC_LOSS_CASES.items():
        P_fg, P_bg, labels = G.mc_loss_inputs(name)
        for t in P_fg + P_bg:
            t.requires_grad_(True)
        nc = case["num_class"]
        ce, dice, bce = nn.CrossEntropyLoss(), DiceLoss(nc), nn.BCEWithLogitsLoss()
        bg_mask = to_inv(labels, nc).float()          # EMCAD/trainer.py:99-101
        ss = [x for x in powerset(list(range(4)))]
        loss = 0.0
        for s in ss:                                   # EMCAD/trainer.py:129-140
            if s == []:
                continue
            iout, ibg = 0.0, 0.0
            for idx in range(len(s)):
                iout += P_fg[s[idx]]
                ibg += P_bg[s[idx]]
            loss += 0.5 * ce(iout, labels.long()) + 0.7 * dice(iout, labels.float(), softmax=True) + 0.3 * bce(ibg, bg_mask)
        loss.backward()
        arrays = {"loss": np.float64(loss.item()), "subsets": np.array([sum(1 << i for i in s) for s in ss if s])}
        for i in range(4):
            arrays[f"dfg{i}"] = _np(P_fg[i].grad)
            arrays[f"dbg{i}"] = _np(P_bg[i].grad)
        _save(name, **arrays)


# --------------------------------------------------------------------------------------
def gen_full():
    for name, case in G.FULL_CASES.items():
        m = R.build_binary(case["model"], **case["kw"])
        sd = synth.synth_state_dict(m.state_dict(), seed=3)
        m.load_state_dict(sd)
        m.train(case["training"])
        x = G.full_input(name)
        with torch.no_grad():
            outs = m(x)
        arrays = {f"out{i}": _np(o, case["stride"]) for i, o in enumerate(outs)}
        arrays["absmax"] = np.array([o.abs().max().item() for o in outs])
        print("   ", name, "absmax", arrays["absmax"])
        _save(name, **arrays)


def gen_tail():
    """Inference tails: the reference's own statements (MyTest_med.py:35-42, EMCAD/utils/utils.py:285-296) run on the final
    upsamples (pranet.py:349-415 / EMCAD networks.py:116-123) of seeded low-res maps."""
    import torch.nn.functional as F
    bin_tail, mc_tail = R.binary_test_tail(), R.multiclass_val_tail()
    for name, case in G.TAIL_BINARY_CASES.items():
        maps, scales = G.tail_binary_inputs(name)
        outs = []
        for b in range(case["B"]):                     # the reference's test loader is batch 1
            ups = [F.interpolate(m[b:b + 1], scale_factor=s, mode="bilinear") for m, s in zip(maps, scales)]
            if case["nmaps"] == 1:                     # V1 rule (MyTest_med.py:98-102): res = res2 only; same statements with zeros added
                ups = ups + [torch.zeros_like(ups[0])] * 3
            outs.append(bin_tail(lambda image: tuple(ups) + (None,) * 4, None, np.zeros(case["gt"], np.float32)))
        _save(name, out=np.stack(outs))
    for name, case in G.TAIL_MC_CASES.items():
        fg, bg = G.tail_mc_inputs(name)
        labels = []
        for b in range(case["B"]):
            P = [F.interpolate(m[b:b + 1], scale_factor=s, mode="bilinear") for m, s in zip(fg + bg, G.TAIL_MC_SCALES * 2)]
            labels.append(mc_tail(lambda inp: P, None).numpy().astype(np.uint8))
        _save(name, labels=np.stack(labels))


GROUPS = {"tail": gen_tail, "sl": gen_structure_loss, "ll": gen_lowres_loss, "head": gen_heads, "mc": gen_multiclass, "mcdec": gen_mc_decoders,
          "mcl": gen_mc_loss, "full": gen_full}

if __name__ == "__main__":
    assert R.available(), "reference not mounted; golden vectors can only be generated in the builder container"
    for g in (sys.argv[1:] or list(GROUPS)):
        print("[golden]", g)
        GROUPS[g]()
