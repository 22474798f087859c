"""The head engine: runs the DSRA head's layers on the pv2 kernels and records its own backward tape.

Why not torch.autograd per layer?  The kernels work on data formats autograd cannot describe (NHWC operand
tensors in bf16 or split tf32 hi/lo planes, raw split-K slabs, gradients that arrive as several slabs to be
summed on load), so the whole head is ONE torch.autograd.Function (`run_head`): its forward executes engine ops
that append backward closures to a tape, its backward replays the tape in reverse.  PyTorch is used for device
memory (caching allocator), the current stream and parameter storage only.

Precision: `bf16` (tensor cores, kind::f16, bf16 operands / fp32 accumulate) or `fp32` (tensor cores, kind::tf32
with every operand split into hi + lo tf32 planes and three products per K step ~ fp32 accuracy).
"""
from __future__ import annotations

import ctypes as C
import math
import os

import torch
import torch.nn as nn

from . import _lib
from .ops import PV2_BF16, PV2_F32, _ratio, _stream

PV2_TF32 = 2
_PRECISION = "auto"

import numpy as np  # noqa: E402

_PACK_DT = np.dtype([("w", "u8"), ("out_f", "u8"), ("out_d", "u8"), ("f_plane", "i8"), ("d_plane", "i8"), ("start", "i8"),
                     ("Cout", "i4"), ("Cin", "i4"), ("KH", "i4"), ("KW", "i4"), ("f_ild", "i4"), ("f_ioff", "i4"), ("f_ooff", "i4"),
                     ("d_ild", "i4"), ("d_ioff", "i4"), ("d_ooff", "i4")], align=True)       # == pv2_pack_desc (88 B)
_UNPACK_DT = np.dtype([("part", "u8"), ("dw", "u8"), ("split_stride", "i8"), ("start", "i8"), ("splits", "i4"), ("Cout", "i4"),
                       ("Cin", "i4"), ("KH", "i4"), ("KW", "i4"), ("Cin_p", "i4"), ("co_off", "i4"), ("pad_", "i4")], align=True)   # == pv2_unpack_desc (64 B)
_BNSEG_DT = np.dtype([("gamma", "u8"), ("beta", "u8"), ("rm", "u8"), ("rv", "u8"), ("nbt", "u8"), ("eps", "f4"), ("momentum", "f4"),
                      ("c_begin", "i4"), ("c_end", "i4")], align=True)                                  # == pv2_bn_seg (56 B)
_BNFUSE_DT = np.dtype([("seg", _BNSEG_DT, (8,)), ("mean", "u8"), ("invstd", "u8"), ("scale", "u8"), ("shift", "u8"), ("part", "u8"),
                       ("counters", "u8"), ("nsegs", "i4"), ("pad_", "i4")], align=True)                 # == pv2_bn_fuse (504 B)
_BNDEFER_DT = np.dtype([("part", "u8"), ("count", "f4"), ("ldc", "i4"), ("c_off", "i4"), ("pad_", "i4"), ("gamma", "u8"), ("beta", "u8"),
                        ("rm", "u8"), ("rv", "u8"), ("nbt", "u8"), ("eps", "f4"), ("momentum", "f4"), ("mean", "u8"), ("invstd", "u8")],
                       align=True)                                                                      # == pv2_bn_defer (88 B)
assert _PACK_DT.itemsize == 88 and _UNPACK_DT.itemsize == 64 and _BNSEG_DT.itemsize == 56 and _BNFUSE_DT.itemsize == 504
assert _BNDEFER_DT.itemsize == 88
_PAR_CTA_BUDGET = int(os.environ.get("PV2_PAR_CTA_BUDGET", "96"))
_PAR_CTA_BUDGET_NARROW = int(os.environ.get("PV2_PAR_CTA_BUDGET_NARROW", "0"))    # sections of 2-3 chains
_BN_ACC_STRIDE, _SUM_STRIDE = 16, 32   # == PV2_BN_ACC_STRIDE (doubles), PV2_SUM_STRIDE (floats) of include/pv2.h
# stream priority of the dependent chains (branch streams; -1 = above the default 0 the weight-gradient companions run at, so a chain's
# next kernel is placed before queued wgrad / unpack CTAs; captured graphs keep it as the kernel nodes' priority).  Head step at
# B = 16 x 352^2: 1.775 ms at priority 0, 1.749 ms at -1 (two runs each, same box).  capture_stream() is what TrainStep / bench_head capture on.
_CHAIN_PRIORITY = int(os.environ.get("PV2_CHAIN_PRIORITY", "-1"))
_BN_FUSED_SLOTS = int(os.environ.get("PV2_BN_FUSED_SLOTS", "296"))
_UNPACK_BATCH = int(os.environ.get("PV2_UNPACK_BATCH", "4"))     # weight-gradient tensors per unpack launch on a companion stream
_ZARENA_FLOATS = 1 << 20   # 4 MB: ~3.5 K conv channels x 32 floats (forward moments) + ~3.5 K x 4 sums x 32 floats (backward) = 0.56 M floats
_COUNTERS = {}     # device -> zero-initialised ticket counters shared by every launch on that device (each launch leaves them zeroed)


def _ticket_counters(device, branch=0):
    """One counter buffer per concurrent branch (stream): kernels that may run at the same time must not share tickets."""
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device(), branch)
    t = _COUNTERS.get(key)
    if t is None:
        t = torch.zeros(4096, dtype=torch.int32, device=device)
        if not torch.cuda.is_current_stream_capturing():     # a buffer born inside a graph capture belongs to that graph's pool
            _COUNTERS[key] = t
    return t


def set_precision(p: str):
    """'bf16' | 'fp32' | 'auto' (bf16 when the backbone features are bf16 / autocast is on, else fp32)."""
    global _PRECISION
    if p not in ("bf16", "fp32", "auto"):
        raise ValueError(p)
    _PRECISION = p


def get_precision() -> str:
    return _PRECISION


def _ptr(t):
    return t.data_ptr() if t is not None else None


def _feat_dt(t: torch.Tensor, who: str) -> int:
    """PV2 dtype code of a feature / gradient tensor handed to the kernels: fp32 or bf16 only.  Anything else (fp16 from
    torch.autocast's default dtype, fp64 ...) must never be reinterpreted as bf16 bits."""
    if t.dtype == torch.float32:
        return PV2_F32
    if t.dtype == torch.bfloat16:
        return PV2_BF16
    raise TypeError(f"pranet_v2_b200 {who}: tensor dtype {t.dtype} is not supported by the pv2 kernels (fp32 or bf16 only); "
                    "use torch.autocast('cuda', dtype=torch.bfloat16) or cast the features")


class Act:
    """Operand-format activation: NHWC, `ld` stored channels (padded), bf16 [N,H,W,ld] or fp32 [planes,N,H,W,ld].
    May be a channel slice [off, off+C) of a wider (concat) buffer."""
    __slots__ = ("t", "N", "H", "W", "C", "ld", "off", "gslabs", "want_grad", "direct")

    def __init__(self, t, N, H, W, C, ld, off=0, want_grad=False):
        self.t, self.N, self.H, self.W, self.C, self.ld, self.off = t, N, H, W, C, ld, off
        self.direct = None        # zero-copy bf16 channels_last input: {"consumers": n, "dx": gradient written by the dgrad epilogue}
        self.gslabs = []          # gradient contributions: (fp32 tensor [M, ld_g], ld_g, off_g)
        self.want_grad = want_grad

    @property
    def M(self):
        return self.N * self.H * self.W



class Raw:
    """Raw conv output: fp32 [splits, M, ld]; after bn_stats slab 0 holds the split sum."""
    __slots__ = ("t", "splits", "M", "ld", "C", "dy", "N", "H", "W", "stats", "stat_bns", "defer")

    def __init__(self, t, splits, N, H, W, ld, C):
        self.t, self.splits, self.N, self.H, self.W, self.ld, self.C = t, splits, N, H, W, ld, C
        self.M = N * H * W
        self.dy = None            # operand-format gradient w.r.t. this raw output (set by the apply backward)
        self.stats = None         # fp32 [4][C]: batch mean, invstd, scale, shift of every channel (training BN, produced with the conv)
        self.stat_bns = {}        # channel offset -> BatchNorm module whose statistics `stats` holds there
        self.defer = None         # persistent conv kernel: {"part", "nparts", "pending": offsets whose per-CTA partial rows are not folded yet}


class Map:
    """Small fp32 NCHW map (head logits at feature resolution) with gradient accumulation."""
    __slots__ = ("t", "grads")

    def __init__(self, t):
        self.t, self.grads = t, []

    def grad(self):
        if not self.grads:
            return None
        g = self.grads[0]
        for h in self.grads[1:]:
            g = g + h
        return g


_SIDE_STREAMS = {}   # device index -> list of side streams for the parallel sections of a head
MAX_BRANCHES = 12


def _side_streams(device):
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _SIDE_STREAMS:
        _SIDE_STREAMS[idx] = [torch.cuda.Stream(device=device, priority=_CHAIN_PRIORITY) for _ in range(MAX_BRANCHES)]
    return _SIDE_STREAMS[idx]


_WGRAD_STREAMS = {}  # device index -> one weight-gradient stream per branch (index 0 = main)


def _wgrad_streams(device):
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _WGRAD_STREAMS:
        _WGRAD_STREAMS[idx] = [torch.cuda.Stream(device=device) for _ in range(MAX_BRANCHES + 1)]
    return _WGRAD_STREAMS[idx]


def capture_stream(device=None):
    """A stream of the chains' priority to capture a step on (torch.cuda.graph(g, stream=...)): the main chain then outranks the
    weight-gradient companions like the branch streams do."""
    return torch.cuda.Stream(device=device, priority=_CHAIN_PRIORITY)


def streams_enabled() -> bool:
    import os
    return os.environ.get("PV2_STREAMS", "1") != "0"


class _Tape(list):
    """Backward tape: ('op', closure, branch) entries plus ('fork', n) / ('join', n) markers of the parallel sections."""

    def __init__(self, eng):
        super().__init__()
        self.eng = eng

    def append(self, fn):
        list.append(self, ("op", fn, self.eng.cur))

    def mark(self, kind, n):
        list.append(self, (kind, n, 0))


class _Branch:
    def __init__(self, eng, i):
        self.eng, self.i, self.ctx = eng, i, None

    def __enter__(self):
        self.eng.cur = self.i
        if self.i > 0:
            self.ctx = torch.cuda.stream(self.eng.side[self.i - 1])
            self.ctx.__enter__()
        return self

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)
        self.eng.cur = 0
        return False


class Engine:
    """Parallel sections.  The head is many small dependent kernels (most launches are 16-242 CTAs and latency bound), but
    its graph is wide: three pyramid levels, three RFB branches and a reverse-attention stack per level.  `fork(n)` /
    `branch(i)` / `join()` run such independent chains on side streams (inside a CUDA-graph capture they become parallel
    branches of the graph); the backward tape mirrors the structure (a forward join is a backward fork).  Discipline that
    keeps the stream-ordered allocator safe without record_stream: the main stream idles between fork and join, every
    section is joined before the next fork, and every buffer the engine allocates stays referenced until the pass ends."""

    def __init__(self, device, precision: str, training: bool, need_grad: bool, cache: dict = None):
        self.cur = 0              # branch the ops being issued belong to (0 = main stream)
        self.side = _side_streams(device) if streams_enabled() else None
        self._open = 0            # branches of the open section
        # Weight gradients are leaves of the backward graph: nothing waits for them before the optimizer.  Each branch hands its
        # wgrad GEMMs to a companion stream, so the chain dgrad -> BN backward -> dgrad ... never queues behind them and they
        # fill whatever SMs the chains leave idle; backward() joins the companions before the gradients are unpacked.
        self.wside = _wgrad_streams(device) if (streams_enabled() and os.environ.get("PV2_WGRAD_STREAMS", "1") != "0") else None
        self._wused = set()
        self._pack_ev, self._late_keys, self._pack_waited = None, set(), set()
        self._keep = []           # every buffer of this pass (see the class docstring)
        self.lib = _lib.load()
        self.cache = cache if cache is not None else {}
        self.groups_seen = []      # conv groups in call order (recorded on the first run, prepacked in one launch afterwards)
        self.packed = {}           # group key -> (w_op fprop layout, w_op dgrad layout or None)
        self.unpack_jobs = []
        self._wjobs = {}               # branch -> unpack jobs of wgrad GEMMs already launched on its companion stream
        self.dev = device
        self.kind = PV2_BF16 if precision == "bf16" else PV2_TF32
        self.nterms = 1 if precision == "bf16" else 3
        self.planes = 1 if precision == "bf16" else 2
        self.cpad = 8 if precision == "bf16" else 4
        self.op_dtype = torch.bfloat16 if precision == "bf16" else torch.float32
        self.training = training
        self.need_grad = need_grad
        self.tape = _Tape(self)
        self.param_grads = {}     # id(param) -> grad tensor
        # One zero-filled arena per pass (ONE memset at the start of the forward, off every chain): the BatchNorm-backward
        # kernels add their block sums into 4*C-float slices of it (pv2_bn_act_bwd `sums_zeroed`).
        self._zarena = torch.zeros(_ZARENA_FLOATS, dtype=torch.float32, device=device) if (need_grad or training) else None
        self._zoff = 0

    # ---- parallel sections -----------------------------------------------------------------------------
    def _set_width(self, n):
        """Tell the conv launcher how many chains run side by side: with four or more, every persistent conv launch is capped at
        96 CTAs (each then walks several tiles through its ring) so that CTAs of the sibling chains are resident next to it
        instead of queueing for the same slots; a lone chain keeps the whole machine (pv2_conv_set_cta_budget)."""
        self.lib.pv2_conv_set_cta_budget(_PAR_CTA_BUDGET if n >= 4 else (_PAR_CTA_BUDGET_NARROW if n >= 2 else 0))
        # the one-launch BatchNorm backward spins on a grid barrier: its CTAs must all be resident, so the n chains share _BN_FUSED_SLOTS
        # CTA slots (half of the 4 x 148 the device holds at the kernel's 64 registers); below 32 CTAs per launch the two-launch form is used
        share = _BN_FUSED_SLOTS // max(n, 1)
        self.lib.pv2_bn_set_fused_grid(share if share >= 32 else 0)

    def _fan_out(self, n):
        self._set_width(n)
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(main)
        for i in range(n):
            self.side[i].wait_event(ev)

    def _fan_in(self, n):
        self._set_width(0)
        main = torch.cuda.current_stream()
        for i in range(n):
            ev = torch.cuda.Event()
            ev.record(self.side[i])
            main.wait_event(ev)

    def fork(self, n):
        """Open a section of n independent branches (branch indices 1..n); no-op when streams are disabled."""
        if self.side is None:
            return
        if self._open or n > len(self.side):
            raise RuntimeError("internal: nested or oversized parallel section")
        self._open = n
        self._fan_out(n)
        if self.need_grad:
            self.tape.mark("fork", n)

    def branch(self, i):
        """Context of branch i (1-based) of the open section."""
        return _Branch(self, i if (self.side is not None and self._open) else 0)

    def join(self):
        if self.side is None or not self._open:
            return
        self._fan_in(self._open)
        if self.need_grad:
            self.tape.mark("join", self._open)
        self._open = 0

    # ---- allocation ---------------------------------------------------------------------------------
    def pad(self, c):
        return (c + self.cpad - 1) // self.cpad * self.cpad

    def new_act(self, N, H, W, C, ld=None, zero_pad=True):
        ld = ld or self.pad(C)
        shape = (N, H, W, ld) if self.planes == 1 else (self.planes, N, H, W, ld)
        t = torch.zeros(shape, dtype=self.op_dtype, device=self.dev) if (zero_pad and ld != C) else torch.empty(shape, dtype=self.op_dtype, device=self.dev)
        self._keep.append(t)
        return Act(t, N, H, W, C, ld, 0, self.need_grad)

    def plane_stride(self, a: Act):
        return a.N * a.H * a.W * a.ld

    def _es(self):
        return 2 if self.kind == PV2_BF16 else 4

    def _act_ptr(self, a: Act):
        return a.t.data_ptr() + a.off * self._es()

    def f32(self, *shape, zero=False):
        t = (torch.zeros if zero else torch.empty)(shape, dtype=torch.float32, device=self.dev)
        self._keep.append(t)
        return t

    def _tile_counters(self, N, H, W, splits):
        """Zero-initialised ticket counters for one split-K conv launch (one per 128-pixel tile x N tile), taken from this pass's zero
        arena: no state survives a pass, nothing is shared between launches.  None when the launch is not split."""
        if splits <= 1:
            return None
        return self.zeros_small(2 * ((N * H * W + 127) // 128)).data_ptr()

    def zeros_small(self, n):
        """n zero-initialised floats (128-byte aligned) from the per-pass arena; a fresh torch.zeros when the arena is exhausted."""
        n4 = (n + 31) // 32 * 32
        if self._zarena is None or self._zoff + n4 > self._zarena.numel():
            return self.f32(n4, zero=True)
        t = self._zarena[self._zoff:self._zoff + n4]
        self._zoff += n4
        return t

    def add_param_grad(self, p, g):
        k = id(p)
        self.param_grads[k] = g if k not in self.param_grads else self.param_grads[k] + g

    # ---- inputs ---------------------------------------------------------------------------------------
    def from_nchw(self, x: torch.Tensor, grad_sink=None):
        """Backbone feature (NCHW fp32 / bf16, any memory format) -> operand tensor.  A bf16 channels_last tensor with
        C % 8 == 0 is used in place (it already IS the operand layout).  grad_sink(dx) receives the input gradient."""
        N, Cc, H, W = x.shape
        if (self.kind == PV2_BF16 and x.dtype == torch.bfloat16 and Cc % 8 == 0
                and x.is_contiguous(memory_format=torch.channels_last) and x.data_ptr() % 16 == 0):
            a = Act(x.permute(0, 2, 3, 1), N, H, W, Cc, Cc, 0, self.need_grad)
            if self.need_grad and grad_sink is not None and not x.is_contiguous():
                a.direct = {"consumers": 0, "dx": None}
        else:
            xc = x.contiguous()
            a = self.new_act(N, H, W, Cc)
            _lib.check(self.lib.pv2_pack_nchw(xc.data_ptr(), _feat_dt(xc, "from_nchw"), a.t.data_ptr(),
                                              self.plane_stride(a), self.planes, self.kind, N, Cc, H * W, a.ld, 0, _stream()), "pv2_pack_nchw")
        if self.need_grad and grad_sink is not None:
            cl = x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous()

            def bwd():
                if a.direct is not None and a.direct["dx"] is not None and not a.gslabs:
                    grad_sink(a.direct["dx"])       # the dgrad GEMM's epilogue already wrote the channels_last bf16 gradient
                    return
                if not a.gslabs:
                    return
                dx = torch.empty((N, Cc, H, W), dtype=x.dtype, device=self.dev,
                                 memory_format=torch.channels_last if cl else torch.contiguous_format)
                self._unpack_slabs(a.gslabs, dx, N, Cc, H * W, cl)
                if a.direct is not None and a.direct["dx"] is not None:
                    dx = dx + a.direct["dx"]
                grad_sink(dx)
            self.tape.append(bwd)
        return a

    def _slab_arrays(self, slabs):
        pp, k1 = _lib.ptr_array([s[0] for s in slabs])
        lds, k2 = _lib.int_array([s[1] for s in slabs])
        offs, k3 = _lib.int_array([s[2] for s in slabs])
        return pp, lds, offs, (k1, k2, k3)

    def _unpack_slabs(self, slabs, dx, N, Cc, HW, channels_last=False):
        pp, lds, offs, keep = self._slab_arrays(slabs)
        _lib.check(self.lib.pv2_unpack_to_nchw(pp, lds, offs, len(slabs), dx.data_ptr(), _feat_dt(dx, "_unpack_slabs"),
                                               N, Cc, HW, int(channels_last), _stream()), "pv2_unpack_to_nchw")

    # ---- weights: every conv group of the head packed by ONE multi-tensor launch ------------------------------------------
    @staticmethod
    def _gkey(convs):
        return tuple(id(c) for c in convs)

    def _alloc_packed(self, convs):
        c0 = convs[0]
        KH, KW = c0.kernel_size
        taps, Cin, Cout = KH * KW, c0.in_channels, sum(c.out_channels for c in convs)
        Cin_p, Cout_p = self.pad(Cin), self.pad(Cout)
        fshape = (Cout, taps, Cin_p) if self.planes == 1 else (self.planes, Cout, taps, Cin_p)
        w_f = (torch.zeros if Cin_p != Cin else torch.empty)(fshape, dtype=self.op_dtype, device=self.dev)
        w_d = None
        if self.need_grad:
            dshape = (Cin, taps, Cout_p) if self.planes == 1 else (self.planes, Cin, taps, Cout_p)
            w_d = (torch.zeros if Cout_p != Cout else torch.empty)(dshape, dtype=self.op_dtype, device=self.dev)
        return w_f, w_d, (KH, KW, taps, Cin, Cout, Cin_p, Cout_p)

    def _pack_groups(self, groups, stream):
        descs, start = [], 0
        for convs in groups:
            w_f, w_d, (KH, KW, taps, Cin, Cout, Cin_p, Cout_p) = self._alloc_packed(convs)
            self.packed[self._gkey(convs)] = (w_f, w_d)
            o = 0
            for cv in convs:
                if not cv.weight.is_contiguous():
                    raise RuntimeError("pv2 conv engine: conv weights must be dense OIHW (these have channels_last strides); "
                                       "convert only the backbone to channels_last, not the DSRA head")
                descs.append((cv.weight.data_ptr(), w_f.data_ptr(), w_d.data_ptr() if w_d is not None else 0,
                              Cout * taps * Cin_p, Cin * taps * Cout_p, start,
                              cv.out_channels, Cin, KH, KW, Cin_p, 0, o, Cout_p, o, 0))
                start += cv.weight.numel()
                o += cv.out_channels
        if descs:
            arr = np.array(descs, dtype=_PACK_DT)
            _lib.check(self.lib.pv2_weight_pack_multi(arr.ctypes.data, len(descs), self.planes, self.kind, stream), "pv2_weight_pack_multi")

    def prepack(self, groups, early=0):
        """Pack the weights of all `groups` (lists of nn.Conv2d fused along Cout) in both layouts with one launch per
        40 tensors; called at the start of a run once the group list of this head is known.  The first `early` groups (the
        level GEMMs that open the head) are packed on the current stream; the rest -- most of the bytes: the 5x5 stacks -- on a
        companion stream while those GEMMs already run, and every stream waits for them before its first other conv."""
        groups = list(groups)
        if self.wside is None or early <= 0 or early >= len(groups):
            return self._pack_groups(groups, _stream())
        # buffers of the late groups are allocated here (current stream) and filled on the companion stream
        self._pack_groups(groups[:early], _stream())
        cur = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(cur)
        ws = self.wside[0]
        ws.wait_event(ev)
        with torch.cuda.stream(ws):
            self._pack_groups(groups[early:], ws.cuda_stream)
            self._pack_ev = torch.cuda.Event()
            self._pack_ev.record(ws)
        self._late_keys = {self._gkey(g) for g in groups[early:]}
        self._pack_waited = set()

    def _await_pack(self, key):
        """Order the current stream after the companion-stream pack if `key`'s weights were packed there."""
        if self._pack_ev is None or key not in self._late_keys:
            return
        cur = torch.cuda.current_stream()
        if cur.cuda_stream not in self._pack_waited:
            cur.wait_event(self._pack_ev)
            self._pack_waited.add(cur.cuda_stream)

    def _unpack(self, jobs, stream):
        """[split][Cout][tap][Cin_p] partials -> OIHW weight gradients, one launch per 56 tensors."""
        if not jobs:
            return
        descs, start = [], 0
        for (part, split_stride, splits, dw_, Cout, Cin, KH, KW, Cin_p, co_off) in jobs:
            descs.append((part.data_ptr(), dw_.data_ptr(), split_stride, start, splits, Cout, Cin, KH, KW, Cin_p, co_off, 0))
            start += dw_.numel()
        arr = np.array(descs, dtype=_UNPACK_DT)
        _lib.check(self.lib.pv2_wgrad_unpack_multi(arr.ctypes.data, len(descs), stream), "pv2_wgrad_unpack_multi")

    def flush_unpack(self):
        """Weight gradients still waiting for their unpack (single-stream mode): all of them in one multi-tensor launch."""
        self._unpack(self.unpack_jobs, _stream())
        self.unpack_jobs = []

    # ---- convolution (optionally several convs of identical geometry fused along Cout) ---------------------
    def _bn_fuse(self, convs, bns, M, Cout):
        """pv2_bn_fuse descriptor (host struct, passed by value into the kernels) for the BatchNorms that follow `convs`."""
        stats = self.f32(4, Cout)
        if self.lib.pv2_conv_fuses_bn_stats(1, 0) == 2:
            ws = self.zeros_small(2 * _BN_ACC_STRIDE * Cout)   # persistent kernel: two DOUBLE accumulators per channel, one 128-byte line each, zero on entry
        else:
            ws = self.f32(self.lib.pv2_bn_fuse_workspace_floats(M, Cout))
        cnt = _ticket_counters(self.dev, self.cur)
        d = np.zeros(1, dtype=_BNFUSE_DT)
        o, n, seg_of = 0, 0, {}
        for cv, bn in zip(convs, bns):
            if bn is not None:
                track = bn.track_running_stats and bn.running_mean is not None
                d["seg"][0, n] = (_ptr(bn.weight) or 0, _ptr(bn.bias) or 0, bn.running_mean.data_ptr() if track else 0,
                                  bn.running_var.data_ptr() if track else 0, bn.num_batches_tracked.data_ptr() if track else 0,
                                  float(bn.eps), 0.1 if bn.momentum is None else float(bn.momentum), o, o + cv.out_channels)
                seg_of[o] = bn
                n += 1
            o += cv.out_channels
        if n > 8:
            raise ValueError("pv2 conv engine: at most 8 BatchNorm modules per fused conv group")
        base = stats.data_ptr()
        d["mean"], d["invstd"], d["scale"], d["shift"] = base, base + 4 * Cout, base + 8 * Cout, base + 12 * Cout
        d["part"], d["counters"], d["nsegs"] = ws.data_ptr(), cnt.data_ptr(), n
        return d, stats, seg_of, (ws, cnt)

    def conv(self, x: Act, convs, bns=None, out_nchw_bias=False):
        """x -> Raw [splits, M, ld] (or, with out_nchw_bias, a biased fp32 NCHW Map straight from the epilogue).
        `bns` (one BatchNorm2d or None per conv): in training mode their batch statistics are produced by the conv
        launch itself (epilogue reduction + ticket fold) or, under split-K, by one grouped statistics pass."""
        convs = list(convs)
        c0 = convs[0]
        KH, KW = c0.kernel_size
        dh, dw = c0.dilation
        for cv in convs:
            if (cv.kernel_size, cv.dilation, cv.in_channels) != (c0.kernel_size, c0.dilation, c0.in_channels) or cv.stride != (1, 1) or cv.groups != 1:
                raise ValueError("pv2 conv engine: fused convs must share geometry; stride 1, groups 1 only")
            if tuple(cv.padding) != (dh * (KH - 1) // 2, dw * (KW - 1) // 2):
                raise ValueError(f"pv2 conv engine: only 'same' padding is supported (got kernel {cv.kernel_size} dil {cv.dilation} pad {cv.padding})")
        if x.off != 0 or x.ld != self.pad(x.C) or x.C != c0.in_channels:
            raise ValueError("pv2 conv engine: conv input must be a whole operand tensor with matching channels")
        Cin, Cin_p = x.C, x.ld
        Cout = sum(cv.out_channels for cv in convs)
        taps = KH * KW
        lib, st = self.lib, _stream()
        if x.direct is not None:
            x.direct["consumers"] += 1
        self.groups_seen.append(convs)
        key = self._gkey(convs)
        if key not in self.packed:      # first run of this head (group list not cached yet): pack this group on its own
            self.prepack([convs])
        self._await_pack(key)
        w_op = self.packed[key][0]
        N, H, W = x.N, x.H, x.W
        if out_nchw_bias:
            assert len(convs) == 1
            out = self.f32(N, Cout, H, W)
            _lib.check(lib.pv2_conv_fwd(self._act_ptr(x), self.plane_stride(x), w_op.data_ptr(), Cout * taps * Cin_p, self.kind, self.nterms,
                                        N, H, W, Cin_p, Cout, KH, KW, dh, dw, 1, out.data_ptr(), 0, 1, _ptr(c0.bias), None, None, st), "pv2_conv_fwd")
            res = Map(out)
        else:
            ld = (Cout + 3) // 4 * 4
            splits = lib.pv2_conv_splits_hint(N, H, W, Cin_p, Cout, KH, KW, self.kind, self.nterms)
            sums = bool(lib.pv2_conv_sums_splits())       # split-K partials are added into ONE zero-initialised slab by the kernel
            raw_t = self.f32(1, N * H * W, ld, zero=True) if (sums and splits > 1) else self.f32(splits, N * H * W, ld)
            fuse = None
            if self.training and bns is not None and any(b is not None for b in bns):
                fuse, stats, seg_of, hold = self._bn_fuse(convs, list(bns), N * H * W, Cout)
            _lib.check(lib.pv2_conv_fwd(self._act_ptr(x), self.plane_stride(x), w_op.data_ptr(), Cout * taps * Cin_p, self.kind, self.nterms,
                                        N, H, W, Cin_p, Cout, KH, KW, dh, dw, 0, raw_t.data_ptr(), ld, splits, None,
                                        fuse.ctypes.data if fuse is not None else None, self._tile_counters(N, H, W, splits), st),
                       "pv2_conv_fwd")
            # the persistent kernel sums the split-K slabs itself (total in slab 0; the other slabs are scratch)
            res = Raw(raw_t, 1 if sums else splits, N, H, W, ld, Cout)
            if self.need_grad and len(convs) > 1:
                # the gradient buffer of a horizontally fused group is filled slice by slice, possibly from several branches:
                # allocate (and zero its padding) here, on the forward stream, not lazily inside one of them
                res.dy = self.new_act(N, H, W, Cout)
            if fuse is not None:
                how = lib.pv2_conv_fuses_bn_stats(splits, 0)
                if not how:      # split-K (or patch tiles): one grouped pass, sums the slabs into slab 0
                    _lib.check(lib.pv2_bn_stats_group(raw_t.data_ptr(), N * H * W * ld, splits, N * H * W, Cout, ld, fuse.ctypes.data, st),
                               "pv2_bn_stats_group")
                elif how == 2:   # per-CTA partial rows: folded by the first kernel that consumes each BatchNorm's channel slice
                    res.defer = {"part": hold[0], "count": float(N * H * W), "pending": set(seg_of.keys())}
                res.stats, res.stat_bns = stats, seg_of
        if self.need_grad:
            self.tape.append(lambda: self._conv_bwd(x, convs, res, KH, KW, dh, dw, Cin, Cin_p, Cout))
        return res

    def _conv_bwd(self, x: Act, convs, res, KH, KW, dh, dw, Cin, Cin_p, Cout):
        lib, st = self.lib, _stream()
        N, H, W, taps = x.N, x.H, x.W, KH * KW
        if isinstance(res, Map):     # biased NCHW head: gradient arrives as an NCHW map
            g = res.grad()
            if g is None:
                return
            cv = convs[0]
            if cv.bias is not None and cv.bias.requires_grad:
                self.add_param_grad(cv.bias, g.sum((0, 2, 3)))
            dy = self.new_act(N, H, W, Cout)
            _lib.check(lib.pv2_pack_nchw(g.contiguous().data_ptr(), PV2_F32, dy.t.data_ptr(), self.plane_stride(dy), self.planes, self.kind,
                                         N, Cout, H * W, dy.ld, 0, st), "pv2_pack_nchw")
        else:
            dy = res.dy
            if dy is None:
                return
        Cout_p = dy.ld
        # wgrad -> parameter gradients
        if any(cv.weight.requires_grad for cv in convs):
            splits = lib.pv2_conv_wgrad_splits_hint(N, H, W, Cin_p, Cout, KH, KW, self.kind)
            part = self.f32(splits, Cout, taps, Cin_p)
            wst, wctx = st, None
            if self.wside is not None:
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream())           # dy (and `part`'s previous life, if any) are ordered before this
                ws = self.wside[self.cur]
                ws.wait_event(ev)
                self._wused.add(self.cur)
                wctx = torch.cuda.stream(ws)
                wctx.__enter__()
                wst = ws.cuda_stream
            jobs, o = [], 0
            for cv in convs:
                if cv.weight.requires_grad:
                    dw_ = torch.empty(cv.weight.shape, dtype=torch.float32, device=self.dev)
                    jobs.append((part, Cout * taps * Cin_p, splits, dw_, cv.out_channels, Cin, KH, KW, Cin_p, o))
                    self.add_param_grad(cv.weight, dw_)      # filled before backward() returns
                o += cv.out_channels
            try:
                _lib.check(lib.pv2_conv_wgrad(self._act_ptr(dy), self.plane_stride(dy), self._act_ptr(x), self.plane_stride(x), self.kind, self.nterms,
                                              N, H, W, Cin_p, Cout_p, Cout, KH, KW, dh, dw, part.data_ptr(), splits, wst), "pv2_conv_wgrad")
                if wctx is not None:
                    # on the companion stream the split sum + OIHW transpose follows its GEMMs in batches of _UNPACK_BATCH layers (one
                    # multi-tensor launch each): off the critical path there, a quarter of the launches of one unpack per layer, and
                    # the end of the backward pass is not a 200 us serial tail of one big unpack over cold partials
                    pend = self._wjobs.setdefault(self.cur, [])
                    pend += jobs
                    if len(pend) >= _UNPACK_BATCH:
                        self._unpack(pend, wst)
                        self._wjobs[self.cur] = []
                else:
                    self.unpack_jobs += jobs                 # single-stream mode: one multi-tensor launch in flush_unpack()
            finally:
                if wctx is not None:
                    wctx.__exit__(None, None, None)
        # dgrad -> raw slabs appended to the input's gradient list
        if x.want_grad and x.gslabs is not None:
            wt = self.packed[self._gkey(convs)][1]           # dgrad layout, packed together with the fprop layout
            ld = (Cin + 3) // 4 * 4
            splits = lib.pv2_conv_splits_hint(N, H, W, Cout_p, Cin, KH, KW, self.kind, self.nterms)
            if x.direct is not None and x.direct["consumers"] == 1 and splits == 1 and x.ld == Cin and x.off == 0:
                # sole consumer of a zero-copy backbone feature: no fp32 slab, no unpack pass -- bf16 channels_last from the epilogue
                dxf = torch.empty((N, Cin, H, W), dtype=torch.bfloat16, device=self.dev, memory_format=torch.channels_last)
                self._keep.append(dxf)
                _lib.check(lib.pv2_conv_fwd(self._act_ptr(dy), self.plane_stride(dy), wt.data_ptr(), Cin * taps * Cout_p, self.kind, self.nterms,
                                            N, H, W, Cout_p, Cin, KH, KW, dh, dw, 2, dxf.data_ptr(), Cin, 1, None, None, None, st), "pv2_conv_fwd(dgrad, bf16)")
                x.direct["dx"] = dxf
                return
            sums = bool(lib.pv2_conv_sums_splits())
            dx = self.f32(1, N * H * W, ld, zero=True) if (sums and splits > 1) else self.f32(splits, N * H * W, ld)
            _lib.check(lib.pv2_conv_fwd(self._act_ptr(dy), self.plane_stride(dy), wt.data_ptr(), Cin * taps * Cout_p, self.kind, self.nterms,
                                        N, H, W, Cout_p, Cin, KH, KW, dh, dw, 0, dx.data_ptr(), ld, splits, None, None,
                                        self._tile_counters(N, H, W, splits), st), "pv2_conv_fwd(dgrad)")
            for s in range(1 if sums else splits):
                x.gslabs.append((dx[s], ld, 0))

    # ---- BN (+ combine, multiplier, relu) -> operand slice or NCHW map -------------------------------------------
    def _affine(self, raw: Raw, off, C, bn, bias):
        """(scale, shift, mean, invstd) for raw[:, off:off+C]; training BN computes batch stats and updates running stats."""
        lib, st = self.lib, _stream()
        scale, shift = self.f32(C), self.f32(C)
        if bn is None:     # plain bias
            scale.fill_(1.0)
            if bias is not None:
                shift.copy_(bias.detach())
            else:
                shift.zero_()
            return scale, shift, None, None, None
        if self.training:
            raise RuntimeError("internal: training-mode BN statistics are computed per Raw, see bn_apply")
        _lib.check(lib.pv2_bn_eval_affine(C, _ptr(bn.weight), _ptr(bn.bias), bn.running_mean.data_ptr(), bn.running_var.data_ptr(),
                                          float(bn.eps), scale.data_ptr(), shift.data_ptr(), st), "pv2_bn_eval_affine")
        return scale, shift, None, None, None

    def _bn_train_stats(self, raw: Raw, off, C, bn):
        if raw.stats is not None and raw.stat_bns.get(off) is bn:      # produced together with the conv
            st4 = raw.stats
            d = None
            if raw.defer is not None and off in raw.defer["pending"]:
                # still per-CTA partial rows: this consumer folds them (pv2_bn_defer) and publishes the four vectors
                raw.defer["pending"].discard(off)
                track = bn.track_running_stats and bn.running_mean is not None
                d = np.zeros(1, dtype=_BNDEFER_DT)
                d[0] = (raw.defer["part"].data_ptr(), raw.defer["count"], raw.C, off, 0, _ptr(bn.weight) or 0, _ptr(bn.bias) or 0,
                        bn.running_mean.data_ptr() if track else 0, bn.running_var.data_ptr() if track else 0,
                        bn.num_batches_tracked.data_ptr() if track else 0, float(bn.eps), 0.1 if bn.momentum is None else float(bn.momentum),
                        st4[0, off:off + C].data_ptr(), st4[1, off:off + C].data_ptr())
            return st4[2, off:off + C], st4[3, off:off + C], st4[0, off:off + C], st4[1, off:off + C], d
        lib, st = self.lib, _stream()
        scale, shift, mean, inv = self.f32(C), self.f32(C), self.f32(C), self.f32(C)
        ws = self.f32(lib.pv2_bn_workspace_floats(raw.M, C))
        mom = 0.1 if bn.momentum is None else float(bn.momentum)
        track = bn.track_running_stats and bn.running_mean is not None
        _lib.check(lib.pv2_bn_stats(raw.t.data_ptr() + off * 4, raw.M * raw.ld, raw.splits, raw.M, C, raw.ld, _ptr(bn.weight), _ptr(bn.bias),
                                    float(bn.eps), mom, _ptr(bn.running_mean) if track else None, _ptr(bn.running_var) if track else None,
                                    _ptr(bn.num_batches_tracked) if track else None, mean.data_ptr(), inv.data_ptr(), scale.data_ptr(),
                                    shift.data_ptr(), ws.data_ptr(), st), "pv2_bn_stats")
        return scale, shift, mean, inv, None

    def bn_apply(self, src1, src2=None, combine=0, mult: Act = None, relu=False, out: Act = None, out_map=False):
        """src = (raw, channel offset, C, bn module or None, bias or None).  Writes into `out` (an Act or Act slice) or
        returns a Map when out_map.  Covers BasicConv2d's BN (+ caller's ReLU), RFB's relu(a + b), the partial
        decoder's products, and biased / BN'd DSRA heads."""
        lib, st = self.lib, _stream()
        raw1, off1, Cc, bn1, bias1 = src1
        stats = []
        for (raw, off, c, bn, bias) in ([src1] + ([src2] if src2 else [])):
            if bn is not None and self.training:
                stats.append(self._bn_train_stats(raw, off, c, bn))      # also folds the split-K slabs into slab 0
            else:
                stats.append(self._affine(raw, off, c, bn, bias))
        (s1, b1, m1, i1, df1) = stats[0]
        (s2, b2, m2, i2, df2) = stats[1] if src2 else (None, None, None, None, None)
        raw2, off2 = (src2[0], src2[1]) if src2 else (None, 0)
        M, HW = raw1.M, raw1.H * raw1.W
        if out_map:
            out_t = self.f32(raw1.N, Cc, raw1.H, raw1.W)
            res = Map(out_t)
            o_ptr, o_plane, o_planes, o_ld, o_off, o_nchw = out_t.data_ptr(), 0, 1, 0, 0, 1
        else:
            if out is None:
                out = self.new_act(raw1.N, raw1.H, raw1.W, Cc)
            res = out
            o_ptr, o_plane, o_planes, o_ld, o_off, o_nchw = out.t.data_ptr(), self.plane_stride(out), self.planes, out.ld, out.off, 0
        # slab 0 already holds the split-K sum when training-mode BN statistics ran over this raw output
        ns1 = 1 if (bn1 is not None and self.training) else raw1.splits
        ns2 = (1 if (src2[3] is not None and self.training) else raw2.splits) if src2 else 1
        fwd_args = (raw1.t.data_ptr(), raw1.ld, off1, ns1, raw1.M * raw1.ld, s1.data_ptr(), b1.data_ptr(),
                    _ptr(raw2.t) if raw2 else None, raw2.ld if raw2 else 0, off2, ns2, (raw2.M * raw2.ld) if raw2 else 0,
                    _ptr(s2), _ptr(b2), combine,
                    (mult.t.data_ptr() if mult is not None else None), self.plane_stride(mult) if mult is not None else 0,
                    self.planes, mult.ld if mult is not None else 0, mult.off if mult is not None else 0, int(relu), M, Cc, HW)
        _lib.check(lib.pv2_act_apply(*fwd_args, o_ptr, o_plane, o_planes, o_ld, o_off, o_nchw,
                                     df1.ctypes.data if df1 is not None else None, df2.ctypes.data if df2 is not None else None,
                                     self.kind, st), "pv2_act_apply")
        if self.need_grad:
            keep = (s1, b1, s2, b2, m1, i1, m2, i2)
            sums0 = self.zeros_small(4 * Cc * _SUM_STRIDE)   # zero until this op's backward adds its block sums there (one 128-byte line per sum)

            def bwd():
                _alive = keep     # fwd_args holds raw device pointers into these tensors: keep them referenced
                if out_map:
                    g = res.grad()
                    if g is None:
                        return
                    g = g.contiguous()
                    dz = (None, None, None, 0, g.data_ptr())
                    hold = g
                else:
                    slabs = res.gslabs if res.gslabs is not None else None
                    if not slabs:
                        return
                    pp, lds, offs, hold = self._slab_arrays(slabs)
                    dz = (pp, lds, offs, len(slabs), None)
                bn_train = 1 if (bn1 is not None and self.training) else 0
                dy1, dy1_ptr = self._dy_slice(raw1, off1)
                dy2, dy2_ptr = self._dy_slice(raw2, off2) if src2 else (None, None)
                dmult = self.f32(M, Cc) if (mult is not None and mult.want_grad) else None
                dg1, db1 = self.f32(Cc), self.f32(Cc)
                dg2, db2 = (self.f32(Cc), self.f32(Cc)) if src2 else (None, None)
                ws = self.f32(lib.pv2_bn_workspace_floats(M, Cc))
                _lib.check(lib.pv2_bn_act_bwd(*fwd_args, *dz, _ptr(m1), _ptr(i1), _ptr(m2), _ptr(i2), bn_train,
                                              _ptr(dmult), Cc, dy1_ptr, self.plane_stride(dy1), self.planes, dy1.ld,
                                              dy2_ptr, self.plane_stride(dy2) if dy2 else 0, self.planes, dy2.ld if dy2 else 0,
                                              dg1.data_ptr(), db1.data_ptr(), _ptr(dg2), _ptr(db2), ws.data_ptr(), sums0.data_ptr(),
                                              self.kind, _stream()),
                           "pv2_bn_act_bwd")
                self._route_bn_grads(src1, dg1, db1, bn_train)
                if src2:
                    self._route_bn_grads(src2, dg2, db2, bn_train)
                if dmult is not None:
                    self._add_grad(mult, dmult, Cc, 0)
            self.tape.append(bwd)
        return res

    def _dy_slice(self, raw: Raw, off):
        """Operand-format gradient buffer of a raw conv output (allocated once; horizontally fused convs fill slices)."""
        if raw.dy is None:
            raw.dy = self.new_act(raw.N, raw.H, raw.W, raw.C)
        return raw.dy, raw.dy.t.data_ptr() + off * self._es()

    def _route_bn_grads(self, src, dgamma, dbeta, bn_train):
        raw, off, Cc, bn, bias = src
        if bn is not None:
            if bn.weight is not None and bn.weight.requires_grad:
                # training BN: the kernel's sum(da*yhat) IS dgamma.  eval BN (affine): z = y*s + b with s = gamma*r, b = beta - rm*s,
                # r = rsqrt(rv+eps), so dgamma = r * (sum(da*y) - rm*sum(da))
                self.add_param_grad(bn.weight, dgamma if bn_train else (dgamma - dbeta * bn.running_mean) * torch.rsqrt(bn.running_var + bn.eps))
            if bn.bias is not None and bn.bias.requires_grad:
                self.add_param_grad(bn.bias, dbeta)
        elif bias is not None and bias.requires_grad:
            self.add_param_grad(bias, dbeta)

    def _add_grad(self, a: Act, g, ld, off):
        if a.gslabs is not None:
            a.gslabs.append((g, ld, off))
        else:
            raise RuntimeError("internal: gradient routed to an operand slice without an owner")

    # ---- concat: allocate the wide buffer first, producers write slices ---------------------------------------------
    def concat_buffer(self, N, H, W, parts):
        """parts = list of channel counts -> (whole Act, [slice Acts]).  Gradients w.r.t. the whole buffer are seen by
        the slices as (slab, ld, off + slice offset)."""
        total = sum(parts)
        whole = self.new_act(N, H, W, total)
        slices, o = [], 0
        for c in parts:
            s = Act(whole.t, N, H, W, c, whole.ld, o, self.need_grad)
            s.gslabs = _SliceGrads(whole, o)
            slices.append(s)
            o += c
        return whole, slices

    # ---- x2 align-corners upsample (partial decoder) ---------------------------------------------------------------------
    def up2(self, a: Act):
        lib = self.lib
        out = self.new_act(a.N, 2 * a.H, 2 * a.W, a.C)
        _lib.check(lib.pv2_up2_nhwc_fwd(a.t.data_ptr(), self.plane_stride(a), self.planes, a.ld, a.off, out.t.data_ptr(), self.plane_stride(out),
                                        self.planes, out.ld, 0, a.N, a.H, a.W, a.C, self.kind, _stream()), "pv2_up2_nhwc_fwd")
        if self.need_grad:
            def bwd():
                if not out.gslabs:
                    return
                pp, lds, offs, hold = self._slab_arrays(list(out.gslabs))
                din = self.f32(a.M, a.C)
                _lib.check(lib.pv2_up2_nhwc_bwd(pp, lds, offs, len(out.gslabs), din.data_ptr(), a.C, a.N, a.H, a.W, a.C, _stream()), "pv2_up2_nhwc_bwd")
                self._add_grad(a, din, a.C, 0)
            self.tape.append(bwd)
        return out

    # ---- small NCHW maps: bilinear resize, DSRA fusion, V1 residual ---------------------------------------------------------
    def resize(self, m: Map, scale_factor=None, size=None, final=False):
        t = m.t
        B, Cc, ih, iw = t.shape
        if size is None:
            oh, ow = int(math.floor(ih * scale_factor)), int(math.floor(iw * scale_factor))
        else:
            oh, ow = size
        rh, rw = _ratio(ih, oh, False, scale_factor), _ratio(iw, ow, False, scale_factor)
        out = self.f32(B, Cc, oh, ow)
        _lib.check(self.lib.pv2_bilinear_fwd(t.data_ptr(), out.data_ptr(), B * Cc, ih, iw, oh, ow, rh, rw, 0, PV2_F32, _stream()), "pv2_bilinear_fwd")
        res = Map(out)
        if self.need_grad:
            def bwd():
                g = res.grad()
                if g is None:
                    return
                g = g.contiguous().float()
                din = self.f32(B, Cc, ih, iw)
                _lib.check(self.lib.pv2_bilinear_bwd(g.data_ptr(), din.data_ptr(), B * Cc, ih, iw, oh, ow, rh, rw, 0, PV2_F32, _stream()), "pv2_bilinear_bwd")
                m.grads.append(din)
            self.tape.append(bwd)
        return res

    def resize_multi(self, maps, scale_factors):
        """The final upsamples of a forward (pranet.py:349-350,370-371,392-393,414-415): up to 8 maps resized to one output
        size by ONE launch, and their 8 gradients pulled back by one launch -- a single 8 MB map cannot fill HBM."""
        B, Cc = maps[0].t.shape[:2]
        geo = []
        for m, sf in zip(maps, scale_factors):
            ih, iw = m.t.shape[-2:]
            geo.append((ih, iw, int(math.floor(ih * sf)), int(math.floor(iw * sf)), _ratio(ih, int(math.floor(ih * sf)), False, sf),
                        _ratio(iw, int(math.floor(iw * sf)), False, sf)))
        oh, ow = geo[0][2], geo[0][3]
        if len(maps) > 8 or any((g[2], g[3]) != (oh, ow) for g in geo) or any(tuple(m.t.shape[:2]) != (B, Cc) for m in maps):
            return [self.resize(m, sf) for m, sf in zip(maps, scale_factors)]
        outs = [self.f32(B, Cc, oh, ow) for _ in maps]
        ihs, k1 = _lib.int_array([g[0] for g in geo])
        iws, k2 = _lib.int_array([g[1] for g in geo])
        rhs = (C.c_float * len(maps))(*[g[4] for g in geo])
        rws = (C.c_float * len(maps))(*[g[5] for g in geo])
        pin, k3 = _lib.ptr_array([m.t for m in maps])
        pout, k4 = _lib.ptr_array(outs)
        _lib.check(self.lib.pv2_bilinear_multi_fwd(pin, pout, ihs, iws, rhs, rws, len(maps), B * Cc, oh, ow, 0, PV2_F32, _stream()), "pv2_bilinear_multi_fwd")
        res = [Map(o) for o in outs]
        if self.need_grad:
            def bwd():
                gs = [r.grad() for r in res]
                live = [i for i, g in enumerate(gs) if g is not None]
                if not live:
                    return
                gl = [gs[i].contiguous().float() for i in live]
                dins = [self.f32(B, Cc, geo[i][0], geo[i][1]) for i in live]
                a1, h1 = _lib.int_array([geo[i][0] for i in live])
                a2, h2 = _lib.int_array([geo[i][1] for i in live])
                b1 = (C.c_float * len(live))(*[geo[i][4] for i in live])
                b2 = (C.c_float * len(live))(*[geo[i][5] for i in live])
                pg, h3 = _lib.ptr_array(gl)
                pd, h4 = _lib.ptr_array(dins)
                _lib.check(self.lib.pv2_bilinear_multi_bwd(pg, pd, a1, a2, b1, b2, len(live), B * Cc, oh, ow, 0, PV2_F32, _stream()), "pv2_bilinear_multi_bwd")
                for i, d in zip(live, dins):
                    maps[i].grads.append(d)
            self.tape.append(bwd)
        return res

    def fuse(self, fg: Map, deep_fg: Map, deep_bg: Map, use_softmax=True, scale_factor=None):
        B, Cc, h, w = fg.t.shape
        dh, dw = deep_fg.t.shape[-2:]
        rh, rw = _ratio(dh, h, False, scale_factor), _ratio(dw, w, False, scale_factor)
        out = self.f32(B, Cc, h, w)
        lib = self.lib
        _lib.check(lib.pv2_dsra_fuse_fwd(fg.t.data_ptr(), deep_fg.t.data_ptr(), deep_bg.t.data_ptr(), out.data_ptr(), B, Cc, h, w, dh, dw, rh, rw,
                                         int(use_softmax), _stream()), "pv2_dsra_fuse_fwd")
        res = Map(out)
        if self.need_grad:
            def bwd():
                g = res.grad()
                if g is None:
                    return
                g = g.contiguous()
                dfg, dd = self.f32(B, Cc, h, w), self.f32(B, Cc, h, w)
                _lib.check(lib.pv2_dsra_fuse_bwd(g.data_ptr(), fg.t.data_ptr(), deep_fg.t.data_ptr(), deep_bg.t.data_ptr(), dfg.data_ptr(), dd.data_ptr(),
                                                 B, Cc, h, w, dh, dw, rh, rw, int(use_softmax), _stream()), "pv2_dsra_fuse_bwd")
                fg.grads.append(dfg)
                if not (Cc == 1 and use_softmax):     # softmax over one channel is constant: exactly zero gradient
                    ddeep = self.f32(B, Cc, dh, dw)
                    _lib.check(lib.pv2_bilinear_bwd(dd.data_ptr(), ddeep.data_ptr(), B * Cc, dh, dw, h, w, rh, rw, 0, PV2_F32, _stream()), "pv2_bilinear_bwd")
                    deep_fg.grads.append(ddeep)
                    deep_bg.grads.append(-ddeep)
            self.tape.append(bwd)
        return res

    def add_maps(self, a: Map, b: Map):
        res = Map(a.t + b.t)
        if self.need_grad:
            def bwd():
                g = res.grad()
                if g is not None:
                    a.grads.append(g)
                    b.grads.append(g)
            self.tape.append(bwd)
        return res

    def ra_v1(self, x: torch.Tensor, crop: Map, grad_sink):
        """V1 reverse attention on an NCHW backbone feature: (1 - sigmoid(crop)) * x -> NCHW tensor (then packed by from_nchw)."""
        lib = self.lib
        xc = x.contiguous()
        B, Cc, h, w = xc.shape
        y = torch.empty_like(xc)
        dt = _feat_dt(xc, "ra_v1")
        _lib.check(lib.pv2_ra_v1_scale_fwd(xc.data_ptr(), crop.t.data_ptr(), y.data_ptr(), B, Cc, h * w, dt, _stream()), "pv2_ra_v1_scale_fwd")

        def sink(dy):
            dy = dy.contiguous()
            dx, dcrop = torch.empty_like(xc), self.f32(B, 1, h, w)
            _lib.check(lib.pv2_ra_v1_scale_bwd(dy.data_ptr(), xc.data_ptr(), crop.t.data_ptr(), dx.data_ptr(), dcrop.data_ptr(), B, Cc, h * w, dt, _stream()),
                       "pv2_ra_v1_scale_bwd")
            crop.grads.append(dcrop)
            grad_sink(dx)
        return y, sink

    def backward(self):
        for kind, x, br in reversed(self.tape):
            if kind == "op":
                with _Branch(self, br if self.side is not None else 0):
                    x()
            elif kind == "join":          # a forward join is a backward fork, and vice versa
                self._fan_out(x)
            else:
                self._fan_in(x)
        if self._wused:            # join the weight-gradient streams
            main = torch.cuda.current_stream()
            for i in sorted(self._wused):
                if self._wjobs.get(i):           # the stream's last (partial) batch of unpacks
                    self._unpack(self._wjobs[i], self.wside[i].cuda_stream)
                ev = torch.cuda.Event()
                ev.record(self.wside[i])
                main.wait_event(ev)
            self._wused = set()
            self._wjobs = {}
        self.flush_unpack()
        self.tape = _Tape(self)
        self._keep = []


class _SliceGrads:
    """gslabs proxy of a concat slice: gradients of the WHOLE buffer, seen at this slice's channel offset."""

    def __init__(self, whole: Act, off: int):
        self.whole, self.off = whole, off

    def __bool__(self):
        return bool(self.whole.gslabs)

    def __len__(self):
        return len(self.whole.gslabs)

    def __iter__(self):
        return iter([(g, ld, o + self.off) for (g, ld, o) in self.whole.gslabs])

    def __getitem__(self, i):
        g, ld, o = self.whole.gslabs[i]
        return (g, ld, o + self.off)

    def append(self, item):
        raise RuntimeError("internal: a concat slice has exactly one consumer (the conv over the whole buffer)")


# ------------------------------------------------------------------------------------------------------------
# autograd boundary
# ------------------------------------------------------------------------------------------------------------
class _HeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, runner, n_inputs, precision, training, cache, *tensors):
        inputs, params = tensors[:n_inputs], tensors[n_inputs:]
        need_grad = any(ctx.needs_input_grad[5:])
        eng = Engine(inputs[0].device, precision, training, need_grad, cache)
        in_grads = [None] * n_inputs
        if cache is not None and "groups" in cache:
            eng.prepack(cache["groups"], cache.get("early_groups", 0))   # every conv weight of this head, both layouts
        outs = runner(eng, inputs, in_grads)
        if cache is not None and "groups" not in cache:
            cache["groups"] = eng.groups_seen
        res = tuple(o.t for o in outs)
        if need_grad:
            # The returned tensor OBJECTS must not be reachable from ctx (output -> grad_fn -> ctx -> output is a cycle that would
            # pin every activation of a forward that is never backpropagated until the cycle collector runs): the tape's Maps keep
            # detached aliases of the same storage (backward closures of the fusion kernels still read them).
            for o in outs:
                o.t = o.t.detach()
            ctx.eng, ctx.in_grads, ctx.params, ctx.outs = eng, in_grads, params, outs
        else:
            eng._keep = []
            ctx.eng = ctx.outs = None
            ctx.in_grads, ctx.params = in_grads, params
        ctx.n_inputs = n_inputs
        return res

    @staticmethod
    def backward(ctx, *gouts):
        eng = ctx.eng
        if eng is None:
            raise RuntimeError("pranet_v2_b200 head: backward called twice (or on a forward that recorded no tape); the engine's tape "
                               "is consumed by the first backward -- retain_graph=True is not supported, run the forward again")
        for o, g in zip(ctx.outs, gouts):
            if g is not None:
                o.grads.append(g)
        eng.backward()
        grads = []
        for i in range(ctx.n_inputs):
            grads.append(ctx.in_grads[i] if ctx.needs_input_grad[5 + i] else None)
        for j, p in enumerate(ctx.params):
            g = eng.param_grads.get(id(p)) if ctx.needs_input_grad[5 + ctx.n_inputs + j] else None
            grads.append(g)
        # drop every other reference to the gradient tensors: autograd's AccumulateGrad adopts a gradient it holds the only
        # reference to and CLONES it otherwise (209 head parameters = 209 serialised 1.4 us copies at the end of the step)
        eng.param_grads = {}
        ctx.eng = ctx.outs = ctx.in_grads = None
        return (None, None, None, None, None, *grads)


def run_head(runner, inputs, params, training, cache=None):
    """runner(engine, inputs, in_grads) -> list of Map.  `inputs` are NCHW feature tensors, `params` the list of
    parameters the runner touches (so autograd can hand their gradients back)."""
    for t in inputs:
        if not t.is_cuda:
            raise RuntimeError("pranet_v2_b200 head is CUDA-only (sm_100a); got a CPU tensor and there is no CPU fallback")
    # fp16 features (torch.autocast's default CUDA dtype, and what the reference's EMCAD / MIST trainers get from
    # torch.cuda.amp.autocast) are converted to bf16 by a differentiable cast -- autograd casts the gradient back; any other
    # dtype is rejected by _feat_dt.  Never reinterpret.
    inputs = [t.to(torch.bfloat16) if t.dtype == torch.float16 else t for t in inputs]
    for t in inputs:
        _feat_dt(t, "run_head")
    prec = _PRECISION
    if prec == "auto":
        autocast_bf16 = torch.is_autocast_enabled() and torch.get_autocast_gpu_dtype() == torch.bfloat16
        prec = "bf16" if (inputs[0].dtype == torch.bfloat16 or autocast_bf16) else "fp32"
    with torch.autocast("cuda", enabled=False):
        return _HeadFn.apply(runner, len(inputs), prec, training, cache, *inputs, *params)
