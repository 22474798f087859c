#!/usr/bin/env python
"""A/B of the tuning knobs of the memory-bound kernels at B = 16 x 352^2 (one process, knobs read per call):
PV2_LOSS_PREFETCH (fused loss forward), PV2_BIL_BAND (final upsamples forward), PV2_BIL_BWD_VARIANT (their backward).
Each variant: 24 launches captured in a CUDA graph over rotating > L2 buffers, CUDA events around the replays."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench_head as BH  # noqa: E402
import pranet_v2_b200 as P  # noqa: E402
from pranet_v2_b200 import synthetic  # noqa: E402
from pranet_v2_b200.ops import PV2_F32, _ratio  # noqa: E402

B, S, dev = 16, 352, "cuda"
hbm = BH.peaks()[0]
lib = P._lib.load()
px = B * S * S
nset = 4
cnt = [0]


def rot():
    cnt[0] += 1
    return cnt[0] % nset


def cur():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


out = []


def report(name, knob, val, nbytes, ms):
    r = {"kernel": name, "knob": knob, "value": val, "us": ms * 1e3, "gbs": nbytes / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": nbytes / (ms * 1e-3) / 1e9 / hbm}
    out.append(r)
    print(f"{r['us']:8.2f} us {r['gbs']:8.1f} GB/s {100 * r['frac_of_hbm_peak']:5.1f}%  {name}  {knob}={val}", flush=True)


# ---- structure loss forward x4
m = synthetic.ellipse_masks(B, S, S, 3).to(dev)
logits = [[torch.randn(B, 1, S, S, device=dev) for _ in range(8)] for _ in range(nset)]
ws_bytes = lib.pv2_structure_loss_workspace_bytes(B, S, S, 4)
ws = torch.empty(ws_bytes // 4, device=dev)
loss = torch.empty(4, device=dev)
packs = [(P._lib.ptr_array(l[:4]), P._lib.ptr_array(l[4:])) for l in logits]


def sl_fwd():
    (pp, _), (pb, _) = packs[rot()]
    P._lib.check(lib.pv2_structure_loss_fwd(pp, pb, m.data_ptr(), None, 4, B, S, S, 0, loss.data_ptr(), ws.data_ptr(), ws_bytes, cur()), "fwd")


ref = None
for v in ("0", "1", "2", "0", "2"):
    os.environ["PV2_LOSS_PREFETCH"] = v
    report("structure_loss fwd x4 (fused)", "PV2_LOSS_PREFETCH", v, px * (4 + 32), BH.timed_graph(sl_fwd))
    cnt[0] = 0
    sl_fwd()
    torch.cuda.synchronize()
    ref = loss.clone() if ref is None else ref
    assert torch.equal(ref, loss), "prefetch changed the loss"
os.environ.pop("PV2_LOSS_PREFETCH")
os.environ["PV2_LOSS_TWO_PASS"] = "1"
report("structure_loss fwd x4 (boundary-weight kernel + streaming forward: 2 launches)", "PV2_LOSS_TWO_PASS", 1, px * (4 + 32), BH.timed_graph(sl_fwd))
os.environ.pop("PV2_LOSS_TWO_PASS")
del logits, packs

# ---- the 8 final maps, forward and backward
scs = (8, 16, 32, 8, 8, 16, 32, 8)
lows = [[torch.randn(B, 1, S // s, S // s, device=dev) for s in scs] for _ in range(nset)]
his = [[torch.randn(B, 1, S, S, device=dev) for _ in scs] for _ in range(nset)]
ihs = (ctypes.c_int * 8)(*[S // s for s in scs])
rr = (ctypes.c_float * 8)(*[_ratio(S // s, S, False, float(s)) for s in scs])
pk = [(P._lib.ptr_array(lows[j]), P._lib.ptr_array(his[j])) for j in range(nset)]
lowpx = sum(B * (S // s) ** 2 for s in scs)


def mf():
    (pl, _), (ph, _) = pk[rot()]
    P._lib.check(lib.pv2_bilinear_multi_fwd(pl, ph, ihs, ihs, rr, rr, 8, B, S, S, 0, PV2_F32, cur()), "bilm")


def mb():
    (pl, _), (ph, _) = pk[rot()]
    P._lib.check(lib.pv2_bilinear_multi_bwd(ph, pl, ihs, ihs, rr, rr, 8, B, S, S, 0, PV2_F32, cur()), "bilmb")


ref = None
for band in (16, 0, 32, 44, 0):        # 0 = the library's own choice
    os.environ["PV2_BIL_BAND"] = str(band)
    report("bilinear fwd, 8 final maps", "PV2_BIL_BAND", band, (8 * px + lowpx) * 4, BH.timed_graph(mf))
    cnt[0] = 0
    mf()
    torch.cuda.synchronize()
    got = torch.stack([t.sum() for t in his[1]])
    ref = got if ref is None else ref
    assert torch.equal(ref, got), "band changed the result"
os.environ.pop("PV2_BIL_BAND", None)
his = [[torch.randn(B, 1, S, S, device=dev) for _ in scs] for _ in range(nset)]
pk = [(P._lib.ptr_array(lows[j]), P._lib.ptr_array(his[j])) for j in range(nset)]
ref = None
for var in (0, 1, 2, 0, 1, 2):
    os.environ["PV2_BIL_BWD_VARIANT"] = str(var)
    report("bilinear bwd, 8 final maps", "PV2_BIL_BWD_VARIANT", var, (8 * px + lowpx) * 4, BH.timed_graph(mb))
    cnt[0] = 0
    mb()
    torch.cuda.synchronize()
    got = torch.cat([t.flatten() for t in lows[1]]).clone()
    ref = got if ref is None else ref
    assert (ref - got).abs().max().item() <= 1e-5 * ref.abs().max().item(), "variant changed the gradient"
os.environ.pop("PV2_BIL_BWD_VARIANT")

# ---- loss from the low-res maps: register budget of the backward
lscs = (8, 16, 32, 8)
lfg = [[torch.randn(B, 1, S // s, S // s, device=dev) * 3 for s in lscs] for _ in range(2)]
dlow = [[torch.empty(B, 1, S // s, S // s, device=dev) for s in lscs] for _ in range(2)]
lws_bytes = lib.pv2_structure_loss_lowres_workspace_bytes(B, S, S, 4)
lws = torch.empty(lws_bytes // 4, device=dev)
lih = (ctypes.c_int * 4)(*[S // s for s in lscs])
lrr = (ctypes.c_float * 4)(*[_ratio(S // s, S, False, float(s)) for s in lscs])
lp = [P._lib.ptr_array(t) for t in lfg + dlow]
gl = torch.ones(4, device=dev)
P._lib.check(lib.pv2_structure_loss_lowres_fwd(lp[0][0], lp[1][0], lih, lih, lrr, lrr, m.data_ptr(), None, 4, B, S, S, loss.data_ptr(), lws.data_ptr(), lws_bytes, cur()), "lowres fwd")


def ll_bwd():
    P._lib.check(lib.pv2_structure_loss_lowres_bwd(lp[0][0], lp[1][0], lih, lih, lrr, lrr, m.data_ptr(), None, gl.data_ptr(), lp[2][0], lp[3][0],
                                                   4, B, S, S, lws.data_ptr(), lws_bytes, cur()), "lowres bwd")


ref = None
for c in (2, 3, 4, 2, 3, 4):
    os.environ["PV2_LOWRES_BWD_CTAS"] = str(c)
    report("structure_loss_lowres bwd x4 + fold", "PV2_LOWRES_BWD_CTAS", c, px * 6, BH.timed_graph(ll_bwd))
    ll_bwd()
    torch.cuda.synchronize()
    got = torch.cat([t.flatten() for t in dlow[0] + dlow[1]]).clone()
    ref = got if ref is None else ref
    assert torch.equal(ref, got), "register budget changed the gradient"
os.environ.pop("PV2_LOWRES_BWD_CTAS")
with open(os.path.join(ROOT, "gpurun_out", "variants.jsonl"), "w") as f:
    for r in out:
        f.write(json.dumps(r) + "\n")
