#!/usr/bin/env python
"""bench_head.py -- BASELINE.json config 5: DSRA head + structure_loss microbenchmark sweep with a roofline report.

    python bench_head.py [--batches 1,4,16,64] [--sizes 256,352,704] [--iters 50] [--precision bf16|fp32]
                         [--backbone res2net|pvt] [--out gpurun_out/head_sweep.jsonl] [--kernels [--kernels-at 64x704]]

For every (batch B, input size S) the head is fed synthetic backbone features relu(randn) of shapes
(B,512,S/8,S/8), (B,1024,S/16,S/16), (B,2048,S/32,S/32) [PVT: 128/320/512] and timed as ONE
CUDA graph: head forward -> 4x structure loss (one fused launch) -> backward through loss and head (feature gradients and
all weight gradients).  Timing: CUDA events around `iters` replays after warm-up; features, activations and gradients
of a replay exceed the 126 MB L2 only at the large end of the sweep -- the small end is launch/latency bound and says so.

`--kernels` adds the per-kernel roofline table: the memory-bound kernels (structure loss fwd/bwd, boundary weight, final
upsample fwd/bwd, V1 reverse-attention scale) in GB/s of ALGORITHMIC bytes against the measured HBM peak, and the conv
GEMMs (fprop of the head's distinct shapes) in TFLOP/s of un-padded 2*M*N*K against the measured bf16 peak.
With torchrun every rank runs the same sweep on its own GPU (independent replicas) and rank 0 reports min/median.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), float(p.get("bf16_tflops", 1590.0)), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, 1590.0, "fallback (B200_PROFILING.md: 6650 GB/s, 1590 TFLOP/s)"


def timed(fn, iters, warm=5):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters   # ms


def timed_graph(fn, n=24, reps=5):
    """Device time of one fn() launch sequence, free of CPU launch overhead: n calls captured in a CUDA graph, replayed."""
    fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    return timed(g.replay, reps, warm=2) / n


def head_flops(model, B, S):
    """Un-padded conv FLOPs of one head forward (2*M*N*K summed over every nn.Conv2d outside the backbone)."""
    import torch.nn as nn
    res = {"backbone.": None}
    total = 0
    # spatial size per conv: derived from the module name (x4-level convs run at S/32, x3 at S/16, x2 at S/8)
    for name, m in model.named_modules():
        if not isinstance(m, nn.Conv2d) or name.startswith(("backbone.", "resnet.", "conv.")):
            continue
        if name.startswith(("rfb4_1", "ra4_")):
            hw = S // 32
        elif name.startswith(("rfb3_1", "ra3_")):
            hw = S // 16
        elif name.startswith(("rfb2_1", "ra2_")):
            hw = S // 8
        elif name.startswith("agg1.conv_upsample1") or name.startswith("agg1.conv_upsample4") or name.startswith("agg1.conv_concat2"):
            hw = S // 16
        else:
            hw = S // 8
        kh, kw = m.kernel_size
        total += 2 * B * hw * hw * m.out_channels * m.in_channels * kh * kw
    return total


def trace_graph(graph, path, B, S, reps=3):
    """Per-kernel device durations INSIDE the graph replay (CUPTI via torch.profiler): what each launch really costs
    when it runs in its place in the step, plus the idle gaps between consecutive kernels."""
    import collections
    import re
    from torch.profiler import ProfilerActivity, profile
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(reps):
            graph.replay()
        torch.cuda.synchronize()
    try:
        prof.export_chrome_trace(path + ".chrome.json")      # per-kernel (stream, start, end): the step's timeline for offline analysis
    except Exception as exc:                                  # noqa: BLE001 -- the aggregate below does not depend on it
        print("chrome trace export failed:", exc)
    evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start],
                 key=lambda e: e.time_range.start)
    agg = collections.defaultdict(lambda: [0, 0.0])
    busy, gaps, last_end = 0.0, 0.0, None
    for e in evs:
        d = e.time_range.end - e.time_range.start
        m = re.search(r"pv2::(?:\(anonymous namespace\)::|<unnamed>::)?(\w+)", e.name)
        name = "pv2::" + m.group(1) if m else ("pv2::slabs_to_nhwc_kernel" if "slabs_to" in e.name else "torch: " + e.name[:60])
        agg[name][0] += 1
        agg[name][1] += d
        busy += d
        if last_end is not None and e.time_range.start > last_end and e.time_range.start - last_end < 200:
            gaps += e.time_range.start - last_end
        last_end = max(last_end or 0, e.time_range.end)
    with open(path, "a") as f:
        f.write(f"# B={B} S={S}: {len(evs) // reps} kernels/replay, sum of kernel durations {busy / reps:.1f} us/replay, idle gaps {gaps / reps:.1f} us/replay\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{v[1] / reps:10.1f} us n={v[0] // reps:5d} avg {v[1] / v[0]:7.2f} us  {k}\n")


def sweep_point(P, model, B, S, iters, chans, dev, lowres=False):
    from pranet_v2_b200 import synthetic
    g = torch.Generator(device="cpu").manual_seed(B * 1000 + S)
    feats = [torch.relu(torch.randn(B, c, S // s, S // s, generator=g)).to(dev).bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
             for c, s in zip(chans, (8, 16, 32))]
    if P.get_precision() == "fp32":
        feats = [f.detach().float().contiguous().requires_grad_(True) for f in feats]
    gt = synthetic.ellipse_masks(B, S, S, 3).to(dev)
    params = model.head_parameters()
    prep_stream = torch.cuda.Stream()

    def step():
        for p in params:
            p.grad = None
        for f in feats:
            f.grad = None
        if lowres:      # head stopped at the low-res maps, final upsamples inside the loss kernels
            outs = model.forward_head(*feats, lowres=True)
            loss = P.structure_loss_lowres([(outs[i], outs[i + 4]) for i in range(4)], model.final_scale_factors(), gt).sum()
        else:
            prepared = P.ops.structure_loss_prepare(gt, prep_stream)      # mask-only boundary weight: a side branch, as in TrainStep
            outs = model.forward_head(*feats)
            loss = P.structure_loss_multi([(outs[i], outs[i + 4]) for i in range(4)], gt, prepared=prepared).sum()
        loss.backward()
        return loss

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    n0 = P._lib.launch_count()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=P.engine.capture_stream()):
        loss = step()
    launches = P._lib.launch_count() - n0
    ms = timed(graph.replay, iters)
    if os.environ.get("PV2_TRACE"):
        trace_graph(graph, os.environ["PV2_TRACE"], B, S)
    ms_eager = timed(step, max(3, iters // 10), warm=1)
    px = B * S * S
    # algorithmic HBM bytes of the full-resolution part of the step (fp32 maps): 8 final maps written (fwd) and their
    # gradients read (bwd) by the upsample kernels, the loss reading 8 maps + mask (fwd) and reading 8 + mask / writing 8 (bwd)
    full_res_bytes = px * 4 * (8 + 8 + (8 + 1) + (8 + 1 + 8))
    return {"B": B, "S": S, "loss_from_lowres": bool(lowres), "ms_graph": ms, "ms_eager": ms_eager, "images_per_s": B / ms * 1e3, "pv2_launches": launches,
            "loss": float(loss), "head_conv_gflop_fwd": head_flops(model, B, S) / 1e9,
            "conv_tflops_fwd_bwd": 3 * head_flops(model, B, S) / (ms * 1e-3) / 1e12,
            "full_res_bytes": full_res_bytes, "full_res_gbs_if_alone": full_res_bytes / (ms * 1e-3) / 1e9}


def kernel_table(P, dev, B, S, hbm, tflops):
    """Per-kernel rooflines at one (B, S): each kernel launched back to back over rotating buffers (> L2), CUDA events."""
    from pranet_v2_b200 import synthetic
    from pranet_v2_b200.ops import PV2_F32, _ratio
    lib = P._lib.load()

    class _St:      # the launching stream is whatever torch's current stream is at call time (a capture stream under timed_graph)
        def __int__(self):
            return torch.cuda.current_stream().cuda_stream
    import ctypes
    def cur():
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rows = []
    px = B * S * S
    nset = max(2, int(400e6 // (px * 4 * 16)) + 1)
    m = synthetic.ellipse_masks(B, S, S, 3).to(dev)
    shape = (B, 1, S, S)
    logits = [[torch.randn(shape, device=dev) for _ in range(8)] for _ in range(nset)]
    grads = [[torch.empty(shape, device=dev) for _ in range(8)] for _ in range(nset)]
    ws_bytes = lib.pv2_structure_loss_workspace_bytes(B, S, S, 4)
    ws = torch.empty(ws_bytes // 4, device=dev)
    loss, gl = torch.empty(4, device=dev), torch.ones(4, device=dev)
    packs = [(P._lib.ptr_array(logits[j][:4]), P._lib.ptr_array(logits[j][4:]), P._lib.ptr_array(grads[j][:4]), P._lib.ptr_array(grads[j][4:])) for j in range(nset)]
    cnt = [0]

    def rot():
        cnt[0] += 1
        return cnt[0] % nset

    def sl_fwd():
        (pp, _), (pb, _), _, _ = packs[rot()]
        P._lib.check(lib.pv2_structure_loss_fwd(pp, pb, m.data_ptr(), None, 4, B, S, S, 0, loss.data_ptr(), ws.data_ptr(), ws_bytes, cur()), "fwd")

    def sl_bwd():
        (pp, _), (pb, _), (dp, _), (dq, _) = packs[rot()]
        P._lib.check(lib.pv2_structure_loss_bwd(pp, pb, m.data_ptr(), None, gl.data_ptr(), dp, dq, 4, B, S, S, 0, ws.data_ptr(), ws_bytes, cur()), "bwd")

    def sl_prep():
        P._lib.check(lib.pv2_structure_loss_prepare(m.data_ptr(), B, S, S, ws.data_ptr(), ws_bytes, cur()), "prep")

    def sl_fwd_prepared():
        (pp, _), (pb, _), _, _ = packs[rot()]
        P._lib.check(lib.pv2_structure_loss_fwd_prepared(pp, pb, m.data_ptr(), None, 4, B, S, S, 0, loss.data_ptr(), ws.data_ptr(), ws_bytes, cur()), "fwdp")

    sl_fwd()
    t = timed_graph(sl_fwd)
    rows.append(("structure_loss fwd x4 (+ boundary weight + finalize)", "hbm", px * (4 + 4 * 8), t))
    sl_prep()
    rows.append(("structure_loss prepare (31x31 boundary weight map of the mask; side branch under the backbone)", "hbm", px * (4 + 2), timed_graph(sl_prep)))
    sl_prep()
    rows.append(("structure_loss fwd x4, prepared (streaming: logits + mask + 2-byte weight map)", "hbm", px * (4 + 2 + 4 * 8), timed_graph(sl_fwd_prepared)))
    t = timed_graph(sl_bwd)
    rows.append(("structure_loss bwd x4", "hbm", px * (4 + 4 * 16), t))
    # the same four losses from the LOW-RES maps (8 f2): final upsamples inside the loss kernels, no full-resolution maps / gradients.
    # Algorithmic bytes: mask read + 16-bit weight map written (fwd) / both read (bwd); these launches are issue bound, the number to
    # compare is their time against [bilinear x8 fwd + loss fwd] and [loss bwd + bilinear x8 bwd] below.
    lscs = (8, 16, 32, 8)
    lfg = [[torch.randn(B, 1, S // s, S // s, device=dev) * 3 for s in lscs] for _ in range(2)]
    dlow = [[torch.empty(B, 1, S // s, S // s, device=dev) for s in lscs] for _ in range(2)]
    lws_bytes = lib.pv2_structure_loss_lowres_workspace_bytes(B, S, S, 4)
    lws = torch.empty(lws_bytes // 4, device=dev)
    masks = [synthetic.ellipse_masks(B, S, S, 3 + j).to(dev) for j in range(nset)]
    lih = (ctypes.c_int * 4)(*[S // s for s in lscs])
    lrr = (ctypes.c_float * 4)(*[_ratio(S // s, S, False, float(s)) for s in lscs])
    lp = [P._lib.ptr_array(t) for t in lfg + dlow]

    def ll_fwd():
        P._lib.check(lib.pv2_structure_loss_lowres_fwd(lp[0][0], lp[1][0], lih, lih, lrr, lrr, masks[rot()].data_ptr(), None, 4, B, S, S,
                                                       loss.data_ptr(), lws.data_ptr(), lws_bytes, cur()), "lowres fwd")

    def ll_bwd():
        P._lib.check(lib.pv2_structure_loss_lowres_bwd(lp[0][0], lp[1][0], lih, lih, lrr, lrr, masks[0].data_ptr(), None, gl.data_ptr(), lp[2][0], lp[3][0],
                                                       4, B, S, S, lws.data_ptr(), lws_bytes, cur()), "lowres bwd")
    ll_fwd()
    rows.append(("structure_loss_lowres fwd x4 (upsamples + boundary weight + loss, one launch)", "hbm", px * (4 + 2), timed_graph(ll_fwd)))
    # the backward reads the weight map / plane sums of the forward that ran last: run it on the mask the backward is given
    ll_fwd_last = lambda: P._lib.check(lib.pv2_structure_loss_lowres_fwd(lp[0][0], lp[1][0], lih, lih, lrr, lrr, masks[0].data_ptr(), None, 4, B, S, S,
                                                                         loss.data_ptr(), lws.data_ptr(), lws_bytes, cur()), "lowres fwd")
    ll_fwd_last()
    rows.append(("structure_loss_lowres bwd x4 (+ fold: 2 launches)", "hbm", px * (4 + 2), timed_graph(ll_bwd)))
    del lfg, dlow, lws, masks
    # final upsamples: x8 of a 44^2 map (two of the 8 maps are x32 / x16; x8 dominates), fp32
    for s in (8, 16, 32):
        h = S // s
        lo = [torch.randn(B, 1, h, h, device=dev) for _ in range(nset * 8)]
        hi = [torch.empty(B, 1, S, S, device=dev) for _ in range(nset * 8)]
        r = _ratio(h, S, False, float(s))

        def up_f():
            j = rot() * 8 % len(lo)
            P._lib.check(lib.pv2_bilinear_fwd(lo[j].data_ptr(), hi[j].data_ptr(), B, h, h, S, S, r, r, 0, PV2_F32, cur()), "bil")

        def up_b():
            j = rot() * 8 % len(lo)
            P._lib.check(lib.pv2_bilinear_bwd(hi[j].data_ptr(), lo[j].data_ptr(), B, h, h, S, S, r, r, 0, PV2_F32, cur()), "bilb")
        rows.append((f"bilinear x{s} fwd (one map)", "hbm", (px + B * h * h) * 4, timed_graph(up_f)))
        rows.append((f"bilinear x{s} bwd (one map)", "hbm", (px + B * h * h) * 4, timed_graph(up_b)))
        del lo, hi
    # the 8 final maps in one launch (x8, x16, x32, x8, twice), as the head issues them
    scs = (8, 16, 32, 8, 8, 16, 32, 8)
    nrot = max(2, nset // 2)
    lows = [[torch.randn(B, 1, S // s, S // s, device=dev) for s in scs] for _ in range(nrot)]
    his = [[torch.empty(B, 1, S, S, device=dev) for _ in scs] for _ in range(nrot)]
    ihs = (ctypes.c_int * 8)(*[S // s for s in scs])
    rr = (ctypes.c_float * 8)(*[_ratio(S // s, S, False, float(s)) for s in scs])
    pk = [(P._lib.ptr_array(lows[j]), P._lib.ptr_array(his[j])) for j in range(nrot)]

    def mf():
        (pl, _), (ph, _) = pk[rot() % nrot]
        P._lib.check(lib.pv2_bilinear_multi_fwd(pl, ph, ihs, ihs, rr, rr, 8, B, S, S, 0, PV2_F32, cur()), "bilm")

    def mb():
        (pl, _), (ph, _) = pk[rot() % nrot]
        P._lib.check(lib.pv2_bilinear_multi_bwd(ph, pl, ihs, ihs, rr, rr, 8, B, S, S, 0, PV2_F32, cur()), "bilmb")
    lowpx = sum(B * (S // s) ** 2 for s in scs)
    rows.append(("bilinear fwd, 8 final maps in one launch", "hbm", (8 * px + lowpx) * 4, timed_graph(mf)))
    rows.append(("bilinear bwd, 8 final maps in one launch", "hbm", (8 * px + lowpx) * 4, timed_graph(mb)))
    del lows, his
    # yardsticks for the rows above: what a plain write-only / read+write stream of the same size reaches on this GPU (stock
    # torch kernels; the "measured HBM peak" is a 2 GiB copy and a 63 MB write-only launch does not get there either)
    big = [torch.empty(8 * px, device=dev) for _ in range(max(3, nrot + 1))]

    def fill():
        big[rot() % len(big)].fill_(1.0)

    def cpy():
        j = rot() % len(big)
        big[j].copy_(big[(j + 1) % len(big)])
    rows.append(("yardstick: torch fill_ of the 8 maps' bytes (write-only stream)", "hbm", 8 * px * 4, timed_graph(fill)))
    rows.append(("yardstick: torch copy_ of the 8 maps' bytes (read + write)", "hbm", 2 * 8 * px * 4, timed_graph(cpy)))
    del big
    # V1 reverse attention scale on the three backbone features (bf16)
    for c, s in ((512, 8), (1024, 16), (2048, 32)):
        h = S // s
        n = max(2, int(300e6 // (B * c * h * h * 4)) + 1)
        xs = [torch.randn(B, c, h, h, device=dev).bfloat16() for _ in range(n)]
        ys = [torch.empty_like(x) for x in xs]
        crop = torch.randn(B, 1, h, h, device=dev)

        def ra():
            j = rot() % n
            P._lib.check(lib.pv2_ra_v1_scale_fwd(xs[j].data_ptr(), crop.data_ptr(), ys[j].data_ptr(), B, c, h * h, 1, cur()), "ra")
        rows.append((f"ra_v1_scale fwd C={c} {h}^2 bf16", "hbm", 2 * B * c * h * h * 2 + B * h * h * 4, timed_graph(ra)))
        del xs, ys
    out = []
    for name, bound, work, ms in rows:
        ach = work / (ms * 1e-3) / 1e9
        out.append({"kernel": name, "bound": bound, "bytes": work, "us": ms * 1e3, "achieved_gbs": ach, "frac_of_hbm_peak": ach / hbm})
    # conv GEMMs (fprop) of the head's distinct big shapes, bf16
    import torch.nn as nn
    for (cin, cout, k, hw, label) in ((512, 224, 1, S // 8, "x2: 6 fused 1x1 (RFB x5 + ra2_conv1)"), (1024, 224, 1, S // 16, "x3: 6 fused 1x1"),
                                      (2048, 416, 1, S // 32, "x4: 6 fused 1x1 (RFB x5 + ra4_conv1)"), (256, 256, 5, S // 32, "ra4_conv2-4 5x5 256->256"),
                                      (96, 96, 3, S // 8, "agg1.conv_concat3 / conv4 3x3 96->96"), (64, 64, 3, S // 8, "ra2_conv2-3 3x3 64->64"),
                                      (32, 32, 3, S // 8, "RFB branch 3x3 32->32 (dil 3)")):
        conv = nn.Conv2d(cin, cout, k, padding=k // 2, bias=False).to(dev)
        eng = P.engine.Engine(torch.device(dev), "bf16", True, False)      # one engine (one zero arena for the statistics accumulators) per shape
        n = max(2, int(300e6 // (B * cin * hw * hw * 2)) + 1)
        acts = [eng.new_act(B, hw, hw, cin) for _ in range(n)]
        for a in acts:
            a.t.normal_()

        bnm = nn.BatchNorm2d(cout).to(dev).train()

        def cv():
            eng.conv(acts[rot() % n], [conv])

        def cv_bn():        # as the head runs it in training: BatchNorm batch statistics produced by the same launch
            eng.conv(acts[rot() % n], [conv], [bnm])
        fl = 2.0 * B * hw * hw * cout * cin * k * k
        for fn, tag in ((cv, "conv_fwd"), (cv_bn, "conv_fwd+bn_stats")):
            fn()
            ms = timed_graph(fn)
            eng._keep.clear()
            # algorithmic HBM bytes: A once (bf16), the packed weights, the fp32 raw output; the roof of a shape is the lower of the two
            hb = B * hw * hw * cin * 2 + cout * cin * k * k * 2 + B * hw * hw * cout * 4
            t_roof = max(fl / (tflops * 1e12), hb / (hbm * 1e9))
            out.append({"kernel": f"{tag} {label} M={B * hw * hw} N={cout} K={cin * k * k}", "bound": "tensor", "flops": fl, "us": ms * 1e3,
                        "achieved_tflops": fl / (ms * 1e-3) / 1e12, "frac_of_bf16_peak": fl / (ms * 1e-3) / 1e12 / tflops,
                        "hbm_bytes": hb, "frac_of_hbm_peak": hb / (ms * 1e-3) / 1e9 / hbm,
                        "roof": "hbm" if hb / (hbm * 1e9) > fl / (tflops * 1e12) else "tensor", "frac_of_own_roof": t_roof / (ms * 1e-3)})
        del acts
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="1,4,16,64")
    ap.add_argument("--sizes", default="256,352,704")
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--backbone", default="res2net", choices=["res2net", "pvt"])
    ap.add_argument("--kernels", action="store_true")
    ap.add_argument("--kernels-at", default="16x352", help="BxS of the per-kernel table, e.g. 64x704 (GEMM-sized conv problems)")
    ap.add_argument("--lowres-loss", default="0", choices=["0", "1", "both"], help="loss from the low-res maps (SURVEY.md 8 f2): off / on / both, one line each")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("bench_head.py needs a CUDA device (there is no CPU path)")
    import pranet_v2_b200 as P
    import torch.distributed as dist
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    P.set_precision(args.precision)
    hbm, tfl, src = peaks()
    torch.manual_seed(0)
    model = (P.PraNet_V2 if args.backbone == "res2net" else P.PVT_PraNet_V2)(num_class=1).to(dev).train()
    chans = (512, 1024, 2048) if args.backbone == "res2net" else (128, 320, 512)
    lines = []
    for S in [int(s) for s in args.sizes.split(",")]:
        for B in [int(b) for b in args.batches.split(",")]:
            if B * S * S > 64 * 704 * 704:
                continue
            for lowres in {"0": (False,), "1": (True,), "both": (False, True)}[args.lowres_loss]:
                r = sweep_point(P, model, B, S, args.iters, chans, dev, lowres)
                if world > 1:
                    t = torch.tensor([r["ms_graph"]], device=dev, dtype=torch.float64)
                    allt = [torch.zeros_like(t) for _ in range(world)]
                    dist.all_gather(allt, t)
                    v = sorted(float(x) for x in allt)
                    r["ms_graph_min_over_gpus"], r["ms_graph_median_over_gpus"], r["n_gpus"] = v[0], v[len(v) // 2], world
                r.update({"precision": args.precision, "backbone": args.backbone, "peaks": src})
                lines.append(r)
                if rank == 0:
                    print(json.dumps(r), flush=True)
                torch.cuda.empty_cache()
    if args.kernels and rank == 0:
        kb, ks = (int(v) for v in args.kernels_at.lower().split("x"))
        for r in kernel_table(P, dev, kb, ks, hbm, tfl):
            r["B"], r["S"] = kb, ks
            r["peaks"] = src
            lines.append(r)
            print(json.dumps(r), flush=True)
    if args.out and rank == 0:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as f:
            for r in lines:
                f.write(json.dumps(r) + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
