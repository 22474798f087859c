"""torch.autograd wrappers over the C ABI (include/pv2.h).  PyTorch supplies device memory, the current
stream and autograd bookkeeping; every computation below happens in libpranetv2_b200.so.

There is no CPU path: tensors that are not on a CUDA device raise.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

PV2_F32, PV2_BF16 = 0, 1


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return PV2_F32
    if t.dtype == torch.bfloat16:
        return PV2_BF16
    raise TypeError(f"pranet_v2_b200 kernels take float32 or bfloat16 tensors, got {t.dtype}")


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("pranet_v2_b200 ops are CUDA-only (sm_100a); got a CPU tensor and there is no CPU fallback")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


# ------------------------------------------------------------------------------------------------
# structure loss
# ------------------------------------------------------------------------------------------------
class PreparedMask:
    """The mask-only part of structure_loss (the 31x31 boundary weight map, MyTrain_med.py:21), computed ahead of the logits by
    `structure_loss_prepare`: holds the contiguous fp32 mask and the loss workspace the weight map lives in."""
    __slots__ = ("mask_fg", "ws", "ws_bytes", "shape", "event")

    def __init__(self, mask_fg, ws, ws_bytes, event):
        self.mask_fg, self.ws, self.ws_bytes, self.shape, self.event = mask_fg, ws, ws_bytes, tuple(mask_fg.shape), event


def structure_loss_prepare(mask_fg: torch.Tensor, stream: "torch.cuda.Stream" = None) -> PreparedMask:
    """Launch the boundary-weight kernel for `mask_fg` (B, C, H, W) now -- on `stream` if given (a side stream that was made to
    wait for the mask), else on the current stream -- and return the handle `structure_loss_multi(..., prepared=handle)` takes.
    A training step calls this before the backbone so that the weight map is off the critical path (train.TrainStep)."""
    _need_cuda(mask_fg)
    lib = _lib.load()
    B, Cc, H, W = mask_fg.shape
    cur = torch.cuda.current_stream()
    m = mask_fg.contiguous().float()
    ws_bytes = lib.pv2_structure_loss_workspace_bytes(B * Cc, H, W, 4)
    ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=m.device)      # allocated on the CURRENT stream: the loss runs there
    ev = None
    if stream is not None and stream != cur:
        stream.wait_stream(cur)
        with torch.cuda.stream(stream):
            _lib.check(lib.pv2_structure_loss_prepare(m.data_ptr(), B * Cc, H, W, ws.data_ptr(), ws_bytes, stream.cuda_stream), "pv2_structure_loss_prepare")
            ev = torch.cuda.Event()
            ev.record(stream)
    else:
        _lib.check(lib.pv2_structure_loss_prepare(m.data_ptr(), B * Cc, H, W, ws.data_ptr(), ws_bytes, _stream()), "pv2_structure_loss_prepare")
    return PreparedMask(m, ws, ws_bytes, ev)


class _StructureLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mask_fg, mask_bg, prepared, *logits):
        lib = _lib.load()
        K = len(logits) // 2
        preds = [t.contiguous() for t in logits[0::2]]
        pred_bgs = [t.contiguous() for t in logits[1::2]]
        B, Cc, H, W = preds[0].shape
        dt = _dt(preds[0])
        for t in preds + pred_bgs:
            if tuple(t.shape) != (B, Cc, H, W) or _dt(t) != dt:
                raise ValueError("structure_loss: all logits must share shape and dtype")
        if tuple(mask_fg.shape) != (B, Cc, H, W):
            raise ValueError(f"structure_loss: mask shape {tuple(mask_fg.shape)} != logits shape {(B, Cc, H, W)}")
        mask_fg = mask_fg.contiguous().float()
        mask_bg = mask_bg.contiguous().float() if mask_bg is not None else None
        planes = B * Cc
        loss = torch.empty(K, dtype=torch.float32, device=mask_fg.device)
        pp, keep1 = _lib.ptr_array(preds)
        pb, keep2 = _lib.ptr_array(pred_bgs)
        if prepared is not None:
            if prepared.shape != (B, Cc, H, W):
                raise ValueError(f"structure_loss: prepared mask shape {prepared.shape} != logits shape {(B, Cc, H, W)}")
            mask_fg, ws, ws_bytes = prepared.mask_fg, prepared.ws, prepared.ws_bytes
            if prepared.event is not None:
                torch.cuda.current_stream().wait_event(prepared.event)
            _lib.check(lib.pv2_structure_loss_fwd_prepared(pp, pb, mask_fg.data_ptr(), mask_bg.data_ptr() if mask_bg is not None else None,
                                                           K, planes, H, W, dt, loss.data_ptr(), ws.data_ptr(), ws_bytes, _stream()),
                       "pv2_structure_loss_fwd_prepared")
        else:
            ws_bytes = lib.pv2_structure_loss_workspace_bytes(planes, H, W, K)
            ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=mask_fg.device)
            _lib.check(lib.pv2_structure_loss_fwd(pp, pb, mask_fg.data_ptr(), mask_bg.data_ptr() if mask_bg is not None else None,
                                                  K, planes, H, W, dt, loss.data_ptr(), ws.data_ptr(), ws_bytes, _stream()),
                       "pv2_structure_loss_fwd")
        ctx.save_for_backward(mask_fg, mask_bg, ws, *preds, *pred_bgs)
        ctx.meta = (K, planes, H, W, dt, ws_bytes)
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        lib = _lib.load()
        K, planes, H, W, dt, ws_bytes = ctx.meta
        mask_fg, mask_bg, ws = ctx.saved_tensors[:3]
        preds = list(ctx.saved_tensors[3:3 + K])
        pred_bgs = list(ctx.saved_tensors[3 + K:3 + 2 * K])
        g = grad_loss.contiguous().float()
        dps = [torch.empty_like(p) for p in preds]
        dqs = [torch.empty_like(p) for p in pred_bgs]
        pp, k1 = _lib.ptr_array(preds)
        pb, k2 = _lib.ptr_array(pred_bgs)
        dp, k3 = _lib.ptr_array(dps)
        dq, k4 = _lib.ptr_array(dqs)
        _lib.check(lib.pv2_structure_loss_bwd(pp, pb, mask_fg.data_ptr(), mask_bg.data_ptr() if mask_bg is not None else None,
                                              g.data_ptr(), dp, dq, K, planes, H, W, dt, ws.data_ptr(), ws_bytes, _stream()),
                   "pv2_structure_loss_bwd")
        grads = []
        for a, b in zip(dps, dqs):
            grads += [a, b]
        return (None, None, None, *grads)


def structure_loss_multi(pairs, mask_fg, mask_bg=None, prepared: PreparedMask = None):
    """`pairs` = [(pred, pred_bg), ...] (1..4 of them) supervised by ONE mask -> tensor of len(pairs) losses.
    One forward launch (+finalize) and one backward launch for all scales; the 31x31 boundary weight is
    computed once per tile instead of once per call (the reference recomputes it 4x, MyTrain_med.py:78-81).
    `prepared` = structure_loss_prepare(mask_fg): the weight map was computed ahead of time and the forward is a pure stream."""
    flat = []
    for p, q in pairs:
        flat += [p, q]
    _need_cuda(mask_fg, mask_bg, *flat)
    if not 1 <= len(pairs) <= 4:
        raise ValueError("structure_loss_multi takes 1..4 (pred, pred_bg) pairs")
    return _StructureLossFn.apply(mask_fg, mask_bg, prepared, *flat)


def structure_loss(pred, pred_bg, mask_fg, mask_bg=None):
    """Drop-in for binary_seg/MyTrain_med.py:19 `structure_loss(pred, pred_bg, mask_fg, mask_bg)`.
    mask_bg=None means 1 - mask_fg (what the reference's training loop passes, MyTrain_med.py:74) and
    saves reading a second mask."""
    return structure_loss_multi([(pred, pred_bg)], mask_fg, mask_bg)[0]


class _StructureLossLowresFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mask_fg, mask_bg, geo, *maps):
        lib = _lib.load()
        K = len(maps) // 2
        fgs = [t.contiguous().float() for t in maps[:K]]
        bgs = [t.contiguous().float() for t in maps[K:]]
        hs, ws_, rhs, rws, (H, W) = geo
        B, Cc = fgs[0].shape[:2]
        mask_fg = mask_fg.contiguous().float()
        mask_bg = mask_bg.contiguous().float() if mask_bg is not None else None
        planes = B * Cc
        ws_bytes = lib.pv2_structure_loss_lowres_workspace_bytes(planes, H, W, K)
        ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=mask_fg.device)
        loss = torch.empty(K, dtype=torch.float32, device=mask_fg.device)
        ctx.save_for_backward(mask_fg, mask_bg, ws, *fgs, *bgs)
        ctx.meta = (K, planes, H, W, ws_bytes, hs, ws_, rhs, rws)
        pf, k1 = _lib.ptr_array(fgs)
        pb, k2 = _lib.ptr_array(bgs)
        ph, k3 = _lib.int_array(hs)
        pw, k4 = _lib.int_array(ws_)
        prh, k5 = _lib.float_array(rhs)
        prw, k6 = _lib.float_array(rws)
        _lib.check(lib.pv2_structure_loss_lowres_fwd(pf, pb, ph, pw, prh, prw, mask_fg.data_ptr(),
                                                     mask_bg.data_ptr() if mask_bg is not None else None, K, planes, H, W,
                                                     loss.data_ptr(), ws.data_ptr(), ws_bytes, _stream()), "pv2_structure_loss_lowres_fwd")
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        lib = _lib.load()
        K, planes, H, W, ws_bytes, hs, ws_, rhs, rws = ctx.meta
        mask_fg, mask_bg, ws = ctx.saved_tensors[:3]
        fgs, bgs = list(ctx.saved_tensors[3:3 + K]), list(ctx.saved_tensors[3 + K:3 + 2 * K])
        g = grad_loss.contiguous().float()
        dfg = [torch.empty_like(t) for t in fgs]
        dbg = [torch.empty_like(t) for t in bgs]
        pf, k1 = _lib.ptr_array(fgs)
        pb, k2 = _lib.ptr_array(bgs)
        df, k3 = _lib.ptr_array(dfg)
        db, k4 = _lib.ptr_array(dbg)
        ph, k5 = _lib.int_array(hs)
        pw, k6 = _lib.int_array(ws_)
        prh, k7 = _lib.float_array(rhs)
        prw, k8 = _lib.float_array(rws)
        _lib.check(lib.pv2_structure_loss_lowres_bwd(pf, pb, ph, pw, prh, prw, mask_fg.data_ptr(),
                                                     mask_bg.data_ptr() if mask_bg is not None else None, g.data_ptr(), df, db,
                                                     K, planes, H, W, ws.data_ptr(), ws_bytes, _stream()), "pv2_structure_loss_lowres_bwd")
        return (None, None, None, *dfg, *dbg)


def lowres_loss_supported(scale_factors, mask_fg) -> bool:
    """True when pv2_structure_loss_lowres_* covers this geometry (fp32 masks with W % 4 == 0, every final upsample >= x4)."""
    W = mask_fg.shape[-1]
    return W % 4 == 0 and all(s >= 4 for s in scale_factors) and mask_fg.data_ptr() % 16 == 0


def structure_loss_lowres(pairs, scale_factors, mask_fg, mask_bg=None):
    """The final upsamples of PraNet_V2.forward (pranet.py:349-350,370-371,392-393,414-415) and the structure_loss calls of
    MyTrain_med.py:78-82 in one pass, from the LOW-RES maps `model.forward_features_lowres` returns.

    pairs: [(low_fg_k, low_bg_k), ...] (1..4), each (B, C, h_k, w_k); scale_factors: the scale_factor of each pair's final
    F.interpolate (`model.final_scale_factors()`); mask_fg (B, C, H, W) with H = floor(h_k * s_k).  Returns the tensor of
    len(pairs) losses; equal to structure_loss_multi on the upsampled maps up to fp32 summation order, without the eight
    full-resolution maps or their gradients ever being written.  Geometries the fused kernels do not cover
    (`lowres_loss_supported`) go through interpolate_bilinear + structure_loss_multi -- the same pv2 kernels the module path uses."""
    fgs, bgs = [p for p, _ in pairs], [q for _, q in pairs]
    _need_cuda(mask_fg, mask_bg, *fgs, *bgs)
    if not 1 <= len(pairs) <= 4 or len(scale_factors) != len(pairs):
        raise ValueError("structure_loss_lowres takes 1..4 (low_fg, low_bg) pairs and one scale factor per pair")
    for a, b in pairs:
        if a.shape != b.shape or a.shape[:2] != mask_fg.shape[:2]:
            raise ValueError("structure_loss_lowres: foreground / background maps must pair up and share (B, C) with the mask")
    hs, ws, rhs, rws, size = _tail_geometry(fgs, scale_factors)
    if tuple(mask_fg.shape[-2:]) != size:
        raise ValueError(f"structure_loss_lowres: maps upsample to {size}, mask is {tuple(mask_fg.shape[-2:])}")
    if not lowres_loss_supported(scale_factors, mask_fg):
        up = [(interpolate_bilinear(a.float(), scale_factor=s), interpolate_bilinear(b.float(), scale_factor=s))
              for (a, b), s in zip(pairs, scale_factors)]
        return structure_loss_multi(up, mask_fg, mask_bg)
    return _StructureLossLowresFn.apply(mask_fg, mask_bg, (hs, ws, rhs, rws, size), *fgs, *bgs)


# ------------------------------------------------------------------------------------------------
# bilinear resize
# ------------------------------------------------------------------------------------------------
def _ratio(in_size: int, out_size: int, align_corners: bool, scale) -> float:
    """ATen's area_pixel_compute_scale<float> (the ratio is rounded to fp32 exactly as ATen does)."""
    if align_corners:
        return float(np.float32(in_size - 1) / np.float32(out_size - 1)) if out_size > 1 else 0.0
    if scale is not None and scale > 0:
        return float(np.float32(1.0 / scale))
    return float(np.float32(in_size) / np.float32(out_size))


class _BilinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, oh, ow, rh, rw, ac):
        lib = _lib.load()
        x = x.contiguous()
        B, Cc, ih, iw = x.shape
        out = torch.empty(B, Cc, oh, ow, dtype=x.dtype, device=x.device)
        _lib.check(lib.pv2_bilinear_fwd(x.data_ptr(), out.data_ptr(), B * Cc, ih, iw, oh, ow, rh, rw, int(ac), _dt(x), _stream()),
                   "pv2_bilinear_fwd")
        ctx.meta = (B, Cc, ih, iw, oh, ow, rh, rw, int(ac))
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        B, Cc, ih, iw, oh, ow, rh, rw, ac = ctx.meta
        g = g.contiguous()
        din = torch.empty(B, Cc, ih, iw, dtype=g.dtype, device=g.device)
        _lib.check(lib.pv2_bilinear_bwd(g.data_ptr(), din.data_ptr(), B * Cc, ih, iw, oh, ow, rh, rw, ac, _dt(g), _stream()),
                   "pv2_bilinear_bwd")
        return din, None, None, None, None, None


def interpolate_bilinear(x, size=None, scale_factor=None, align_corners=False):
    """F.interpolate(x, size=|scale_factor=, mode='bilinear', align_corners=) on the pv2 kernels
    (binary_seg/lib/pranet.py:349-415, :93; EMCAD/lib/decoders.py:460-461)."""
    _need_cuda(x)
    ih, iw = x.shape[-2:]
    if size is not None:
        oh, ow = (size, size) if isinstance(size, int) else tuple(size)
        sh = sw = None
    else:
        sh, sw = (scale_factor, scale_factor) if not isinstance(scale_factor, (tuple, list)) else scale_factor
        oh, ow = int(np.floor(ih * sh)), int(np.floor(iw * sw))
    rh, rw = _ratio(ih, oh, align_corners, sh), _ratio(iw, ow, align_corners, sw)
    return _BilinearFn.apply(x, oh, ow, rh, rw, bool(align_corners))


# ------------------------------------------------------------------------------------------------
# DSRA fusion
# ------------------------------------------------------------------------------------------------
class _DsraFuseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fg, deep_fg, deep_bg, use_softmax, rh, rw):
        lib = _lib.load()
        fg, deep_fg, deep_bg = fg.contiguous().float(), deep_fg.contiguous().float(), deep_bg.contiguous().float()
        B, Cc, h, w = fg.shape
        dh, dw = deep_fg.shape[-2:]
        out = torch.empty_like(fg)
        _lib.check(lib.pv2_dsra_fuse_fwd(fg.data_ptr(), deep_fg.data_ptr(), deep_bg.data_ptr(), out.data_ptr(),
                                         B, Cc, h, w, dh, dw, rh, rw, int(use_softmax), _stream()), "pv2_dsra_fuse_fwd")
        ctx.save_for_backward(fg, deep_fg, deep_bg)
        ctx.meta = (B, Cc, h, w, dh, dw, rh, rw, int(use_softmax))
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        fg, deep_fg, deep_bg = ctx.saved_tensors
        B, Cc, h, w, dh, dw, rh, rw, sm = ctx.meta
        g = g.contiguous().float()
        dfg, dd = torch.empty_like(fg), torch.empty_like(fg)
        _lib.check(lib.pv2_dsra_fuse_bwd(g.data_ptr(), fg.data_ptr(), deep_fg.data_ptr(), deep_bg.data_ptr(),
                                         dfg.data_ptr(), dd.data_ptr(), B, Cc, h, w, dh, dw, rh, rw, sm, _stream()),
                   "pv2_dsra_fuse_bwd")
        ddeep = torch.empty_like(deep_fg)
        _lib.check(lib.pv2_bilinear_bwd(dd.data_ptr(), ddeep.data_ptr(), B * Cc, dh, dw, h, w, rh, rw, 0, PV2_F32, _stream()),
                   "pv2_bilinear_bwd")
        return dfg, ddeep, -ddeep, None, None, None


def dsra_fuse(fg, deep_fg, deep_bg, use_softmax=True, scale_factor=None):
    """fg + fg * softmax_c(up(deep_fg) - up(deep_bg)) with the resize of the deeper maps fused in
    (pranet.py:353-368; EMCAD/lib/decoders.py:460-477).  `scale_factor` mirrors the reference call
    form F.interpolate(scale_factor=) (None = the size= form)."""
    _need_cuda(fg, deep_fg, deep_bg)
    h, w = fg.shape[-2:]
    dh, dw = deep_fg.shape[-2:]
    if scale_factor is not None and (int(np.floor(dh * scale_factor)), int(np.floor(dw * scale_factor))) != (h, w):
        raise ValueError(f"dsra_fuse: {dh}x{dw} * {scale_factor} does not give {h}x{w}")
    return _DsraFuseFn.apply(fg, deep_fg, deep_bg, bool(use_softmax),
                             _ratio(dh, h, False, scale_factor), _ratio(dw, w, False, scale_factor))


# ------------------------------------------------------------------------------------------------
# V1 reverse attention
# ------------------------------------------------------------------------------------------------
class _RaV1Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, crop):
        lib = _lib.load()
        x, crop = x.contiguous(), crop.contiguous().float()
        B, Cc, h, w = x.shape
        y = torch.empty_like(x)
        _lib.check(lib.pv2_ra_v1_scale_fwd(x.data_ptr(), crop.data_ptr(), y.data_ptr(), B, Cc, h * w, _dt(x), _stream()),
                   "pv2_ra_v1_scale_fwd")
        ctx.save_for_backward(x, crop)
        return y

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        x, crop = ctx.saved_tensors
        B, Cc, h, w = x.shape
        g = g.contiguous().to(x.dtype)
        dx, dcrop = torch.empty_like(x), torch.empty_like(crop)
        _lib.check(lib.pv2_ra_v1_scale_bwd(g.data_ptr(), x.data_ptr(), crop.data_ptr(), dx.data_ptr(), dcrop.data_ptr(),
                                           B, Cc, h * w, _dt(x), _stream()), "pv2_ra_v1_scale_bwd")
        return dx, dcrop


def ra_v1_scale(x, crop):
    """(1 - sigmoid(crop)).expand(-1, C, -1, -1) * x  (binary_seg/lib/PraNet_Res2Net.py:153-154)."""
    _need_cuda(x, crop)
    if crop.shape[1] != 1 or crop.shape[0] != x.shape[0] or crop.shape[-2:] != x.shape[-2:]:
        raise ValueError(f"ra_v1_scale: crop {tuple(crop.shape)} does not broadcast over x {tuple(x.shape)}")
    return _RaV1Fn.apply(x, crop)


# ------------------------------------------------------------------------------------------------
# multiclass dual-supervision loss
# ------------------------------------------------------------------------------------------------
class _McDualLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, labels, num_classes, mode, lc, *logits):
        lib = _lib.load()
        n = len(logits) // 2
        fgs = [t.contiguous().float() for t in logits[:n]]
        bgs = [t.contiguous().float() for t in logits[n:]]
        B, Cc, H, W = fgs[0].shape
        if Cc != num_classes:
            raise ValueError(f"mc_dual_loss: logits have {Cc} channels, num_classes={num_classes}")
        for t in fgs + bgs:
            if tuple(t.shape) != (B, Cc, H, W):
                raise ValueError("mc_dual_loss: all logits must share one shape")
        if tuple(labels.shape) != (B, H, W):
            raise ValueError(f"mc_dual_loss: labels shape {tuple(labels.shape)} != {(B, H, W)}")
        labels = labels.contiguous().long()
        ws_bytes = lib.pv2_mc_dual_loss_workspace_bytes(B, Cc, H, W)
        ws = torch.empty((ws_bytes + 3) // 4, dtype=torch.float32, device=labels.device)
        loss = torch.empty((), dtype=torch.float32, device=labels.device)
        pf, k1 = _lib.ptr_array(fgs)
        pb, k2 = _lib.ptr_array(bgs)
        _lib.check(lib.pv2_mc_dual_loss_fwd(pf, pb, labels.data_ptr(), n, mode, B, Cc, H, W, lc[0], lc[1], lc[2], loss.data_ptr(),
                                            ws.data_ptr(), ws_bytes, _stream()), "pv2_mc_dual_loss_fwd")
        ctx.save_for_backward(labels, ws, *fgs, *bgs)
        ctx.meta = (n, mode, B, Cc, H, W, lc, ws_bytes)
        return loss

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        n, mode, B, Cc, H, W, lc, ws_bytes = ctx.meta
        labels, ws = ctx.saved_tensors[:2]
        fgs, bgs = list(ctx.saved_tensors[2:2 + n]), list(ctx.saved_tensors[2 + n:2 + 2 * n])
        g = g.contiguous().float()
        dfg = [torch.empty_like(t) for t in fgs]
        dbg = [torch.empty_like(t) for t in bgs]
        pf, k1 = _lib.ptr_array(fgs)
        pb, k2 = _lib.ptr_array(bgs)
        df, k3 = _lib.ptr_array(dfg)
        db, k4 = _lib.ptr_array(dbg)
        _lib.check(lib.pv2_mc_dual_loss_bwd(pf, pb, labels.data_ptr(), g.data_ptr(), df, db, n, mode, B, Cc, H, W, lc[0], lc[1], lc[2],
                                            ws.data_ptr(), ws_bytes, _stream()), "pv2_mc_dual_loss_bwd")
        return (None, None, None, None, *dfg, *dbg)


def mc_dual_loss(P_fg, P_bg, labels, num_classes, lc=(0.5, 0.7, 0.3), supervision="mutation"):
    """The multiclass dual loss block of EMCAD/trainer.py:123-140: P_fg / P_bg are the lists of (up to 4) foreground /
    background logits (B, C, H, W), labels the (B, H, W) class map; `supervision` = 'mutation' (all non-empty subsets,
    the reference default) or 'deep_supervision' (each scale alone).  The inverted one-hot mask of trainer.py:22-29 is
    derived on the device from `labels`."""
    P_fg, P_bg = list(P_fg), list(P_bg)
    _need_cuda(labels, *P_fg, *P_bg)
    if len(P_fg) != len(P_bg) or not 1 <= len(P_fg) <= 4:
        raise ValueError("mc_dual_loss takes 1..4 foreground maps and as many background maps")
    mode = {"mutation": 0, "deep_supervision": 1}[supervision]
    return _McDualLossFn.apply(labels, int(num_classes), mode, tuple(float(v) for v in lc), *P_fg, *P_bg)


# ------------------------------------------------------------------------------------------------
# inference tails from the low-resolution head maps
# ------------------------------------------------------------------------------------------------
def _tail_geometry(maps, scale_factors):
    hs, ws, rhs, rws = [], [], [], []
    size = None
    for t, s in zip(maps, scale_factors):
        h, w = t.shape[-2:]
        out = (int(np.floor(h * s)), int(np.floor(w * s)))
        if size is not None and out != size:
            raise ValueError(f"inference tail: maps upsample to different sizes ({size} vs {out})")
        size = out
        hs.append(h); ws.append(w)
        rhs.append(_ratio(h, out[0], False, s)); rws.append(_ratio(w, out[1], False, s))
    return hs, ws, rhs, rws, size


@torch.no_grad()
def infer_tail_binary(maps, scale_factors, size=None):
    """uint8 saliency maps of binary_seg/MyTest_med.py:35-42 / :104-111 from the LOW-RES foreground maps.

    maps: 1..4 tensors (B, 1, h_k, w_k); scale_factors: the factors the model's final F.interpolate calls use
    (pranet.py:349-415: 8, 16, 32, 8 over sem_downsample); size: the ground-truth (H, W) the reference resizes to
    (None = the model output size).  Returns uint8 (B, H, W); min-max normalisation is per image."""
    maps = [m.contiguous().float() for m in maps]
    _need_cuda(*maps)
    lib = _lib.load()
    B = maps[0].shape[0]
    if any(m.dim() != 4 or m.shape[0] != B or m.shape[1] != 1 for m in maps):
        raise ValueError("infer_tail_binary takes (B, 1, h, w) maps")
    hs, ws, rhs, rws, (SH, SW) = _tail_geometry(maps, scale_factors)
    GH, GW = (SH, SW) if size is None else tuple(int(v) for v in size)
    out = torch.empty(B, GH, GW, dtype=torch.uint8, device=maps[0].device)
    ws_bytes = lib.pv2_infer_tail_workspace_bytes(B)
    wsp = torch.empty((ws_bytes + 3) // 4, dtype=torch.int32, device=maps[0].device)
    pm, k0 = _lib.ptr_array(maps)
    ph, k1 = _lib.int_array(hs)
    pw, k2 = _lib.int_array(ws)
    prh, k3 = _lib.float_array(rhs)
    prw, k4 = _lib.float_array(rws)
    _lib.check(lib.pv2_infer_tail_binary(pm, ph, pw, prh, prw, len(maps), B, SH, SW, GH, GW, _ratio(SH, GH, False, None), _ratio(SW, GW, False, None),
                                         out.data_ptr(), wsp.data_ptr(), ws_bytes, _stream()), "pv2_infer_tail_binary")
    return out


@torch.no_grad()
def infer_tail_argmax(P_fg, P_bg, scale_factors):
    """Label maps of EMCAD/utils/utils.py:261-273 (use_dual): argmax_c sum_k (up(P_fg_k) - up(P_bg_k)) from the LOW-RES
    (B, C, h_k, w_k) stage maps; scale_factors as in EMCAD/lib/networks.py:116-123 (32, 16, 8, 4).  Returns uint8 (B, H, W)."""
    P_fg = [m.contiguous().float() for m in P_fg]
    P_bg = [m.contiguous().float() for m in P_bg]
    _need_cuda(*P_fg, *P_bg)
    lib = _lib.load()
    B, Cc = P_fg[0].shape[:2]
    if len(P_fg) != len(P_bg) or any(a.shape != b.shape or a.shape[:2] != (B, Cc) for a, b in zip(P_fg, P_bg)):
        raise ValueError("infer_tail_argmax: foreground / background maps must pair up and share (B, C)")
    hs, ws, rhs, rws, (H, W) = _tail_geometry(P_fg, scale_factors)
    out = torch.empty(B, H, W, dtype=torch.uint8, device=P_fg[0].device)
    pf, k0 = _lib.ptr_array(P_fg)
    pb, k5 = _lib.ptr_array(P_bg)
    ph, k1 = _lib.int_array(hs)
    pw, k2 = _lib.int_array(ws)
    prh, k3 = _lib.float_array(rhs)
    prw, k4 = _lib.float_array(rws)
    _lib.check(lib.pv2_infer_tail_argmax(pf, pb, ph, pw, prh, prw, len(P_fg), B, Cc, H, W, out.data_ptr(), _stream()), "pv2_infer_tail_argmax")
    return out
