#!/usr/bin/env python
"""A/B of the tuning knobs of the memory-bound kernels at B = 16 x 352^2 (one process, knobs read per call):
PV2_LOSS_TWO_PASS (loss forward as boundary-weight kernel + streaming kernel) and PV2_BIL_BAND (rows per CTA of the final upsamples' forward).  (The knobs whose A/B is recorded in profiles/r1_kernel_knobs_v3.jsonl -- loss-forward L2 prefetch, bilinear-backward
batch x occupancy, low-res-backward register budget -- were settled by that measurement and removed from the kernels.)
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


report("structure_loss fwd x4 (fused)", "PV2_LOSS_TWO_PASS", 0, px * (4 + 32), BH.timed_graph(sl_fwd))
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
report("bilinear bwd, 8 final maps", "-", 0, (8 * px + lowpx) * 4, BH.timed_graph(mb))
with open(os.path.join(ROOT, "gpurun_out", "variants.jsonl"), "w") as f:
    for r in out:
        f.write(json.dumps(r) + "\n")
