#!/usr/bin/env python
"""bench.py -- headline benchmark: PraNet-V2 (Res2Net-50) training step, batch 16 @ 352x352 per GPU,
synthetic data, images/s; plus the roofline of the dominant pv2 kernel and the CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--size S]

N > 1 is launched by torchrun (one rank per GPU, NCCL); weak scaling (per-GPU batch fixed).
Prints ONE JSON line on rank 0.  See DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "train images/sec @352^2 (PraNet-V2 Res2Net-50, DSRA head + structure loss, fwd+bwd+Adam)"
UNIT = "images/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="per-GPU batch")
    ap.add_argument("--size", type=int, default=352)
    ap.add_argument("--cpu-batch", type=int, default=4, help="batch of one CPU-baseline step (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--mode", default="train", choices=["train", "infer"],
                    help="train (default, the headline metric) or infer: batched test-time path, images/s of uint8 saliency maps")
    ap.add_argument("--profile-only", action="store_true",
                    help="run the warm-up + timed steps only (no e2e / roofline / CPU legs) and exit: the command ncu wraps for the launch list")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.proc, self.lines, self.first = index, None, [], 0

    def mark(self):
        """Samples before this call (warm-up) are dropped from the report."""
        self.first = len(self.lines)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in (self.lines[self.first:] or self.lines[-3:]):
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port of the reference's path on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_train_step_factory(batch: int, size: int):
    """One training step of the reference's path on CPU: stock Res2Net-50 backbone (torch CPU) + the
    oracle's functional DSRA head + 4x oracle structure_loss + backward + clamp + Adam, fp32, all host threads."""
    from oracle import dsra_oracle as O
    from oracle import synth, templates
    from pranet_v2_b200.backbones import Res2Net50

    torch.set_num_threads(os.cpu_count() or 1)
    bb = Res2Net50().train()
    sd = synth.synth_state_dict(templates.pranet_head(num_class=1), seed=0)
    params = [v.requires_grad_(True) for k, v in sd.items() if v.dtype.is_floating_point and "running" not in k]
    opt = torch.optim.Adam(list(bb.parameters()) + params, 1e-4)
    x = torch.randn(batch, 3, size, size, generator=torch.Generator().manual_seed(1000))
    gt = synth.ellipse_masks(batch, size, size, 1000)

    def step():
        opt.zero_grad(set_to_none=True)
        _, x2, x3, x4 = bb.pyramid(x)
        outs = O.pranet_v2_head(x2, x3, x4, sd, training=True)
        loss = sum(O.structure_loss(outs[i], outs[i + 4], gt, 1 - gt) for i in range(4))
        loss.backward()
        for p in opt.param_groups[0]["params"]:
            if p.grad is not None:
                p.grad.clamp_(-0.5, 0.5)
        opt.step()
        return float(loss)
    return step


def time_cpu(batch, size, steps, warmup):
    step = cpu_train_step_factory(batch, size)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps


def train_config(args, world):
    """The workload both arms are quoted on (BASELINE.json configs[1])."""
    return {"workload": f"PraNet-V2 Res2Net-50 train step (fwd + 4x structure_loss + bwd + clamp + Adam), per-GPU batch {args.batch} @ {args.size}^2, random init",
            "global_batch": world * args.batch, "parallelism": f"dp{world}", "cuda_graph": not args.no_graph,
            "l2": "inputs rotate over 4 batches; per-step working set (activations+grads) >> 126 MB L2"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ips, per = time_cpu(args.cpu_batch, args.size, args.steps, args.warmup)
    cores = torch.get_num_threads()
    sample = f"{args.steps} timed steps of batch {args.cpu_batch} @ {args.size}^2 (oracle port: stock Res2Net-50 on torch-CPU + oracle head/loss, fp32)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": train_config(args, int(os.environ.get("WORLD_SIZE", "1"))),
        "cpu_baseline": {"value": ips, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    import pranet_v2_b200 as P
    from pranet_v2_b200.train import TrainStep
    from pranet_v2_b200 import synthetic as synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.benchmark = True

    B, S = args.batch, args.size
    torch.manual_seed(0)
    model = P.PraNet_V2(num_class=1)
    if args.mode == "infer":
        return run_infer(args, P, model, dev, world, rank, local)
    ts = TrainStep(model, lr=1e-4, clip=0.5, autocast_backbone=(args.precision == "bf16"), device=dev, use_graph=not args.no_graph)
    g = torch.Generator().manual_seed(1000 + rank)
    # a few distinct batches (> L2 together with the activations; inputs change step to step)
    nbuf = 4
    imgs_h = [torch.randn(B, 3, S, S, generator=g).pin_memory() for _ in range(nbuf)]
    gts_h = [synth.ellipse_masks(B, S, S, 1000 + rank * 17 + i).pin_memory() for i in range(nbuf)]
    imgs_d = [t.to(dev) for t in imgs_h]
    gts_d = [t.to(dev) for t in gts_h]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()            # nvidia-smi needs a few hundred ms before its first sample: start it ahead of the warm-up
    for i in range(args.warmup):
        ts.step_device(imgs_d[i % nbuf], gts_d[i % nbuf])
    barrier()
    if rank == 0:
        sampler.mark()             # only samples taken from here on (the timed region) are reported
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        ts.step_device(imgs_d[i % nbuf], gts_d[i % nbuf])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ts.pv2_launches_per_step * args.steps   # pv2 kernel nodes replayed inside the CUDA graphs of the timed steps
    clocks = sampler.stop() if rank == 0 else None

    if args.profile_only:
        if rank == 0:
            print(json.dumps({"profile_only": True, "ms_per_step": ms / args.steps, "gpu_launches": int(launches), "steps": args.steps}))
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- end-to-end timing: pinned host inputs -> H2D -> step -> loss D2H every step ----
    # every step: its own batch crosses PCIe (prefetched during the previous step's compute, like a data loader) and its loss is read back
    for i in range(2):
        ts.step_host(imgs_h[i % nbuf], gts_h[i % nbuf], next_batch=(imgs_h[(i + 1) % nbuf], gts_h[(i + 1) % nbuf]))
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for i in range(args.steps):
        j = i + 2
        ts.step_host(imgs_h[j % nbuf], gts_h[j % nbuf], next_batch=(imgs_h[(j + 1) % nbuf], gts_h[(j + 1) % nbuf]))
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)

    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()

    # ---- roofline of the dominant pv2 kernel, timed live with CUDA events on the launching stream ----
    roof = None
    if rank == 0:
        roof = roofline_adam(P, dev, ts.bucket.n if hasattr(ts.bucket, "n") else ts.flat.numel())
        roof["other_kernels"] = [roofline_structure_loss(P, dev, B, S)]
        try:
            roof["other_kernels"].append(roofline_bilinear(P, dev, B, S))
        except Exception as exc:      # noqa: BLE001 -- an auxiliary measurement must not take the bench line down with it
            roof["other_kernels"].append({"kernel": "bilinear_fwd/bwd_kernel (8 final maps)", "error": str(exc)[:200]})

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm = float(peaks.get("hbm_gbs", 6650.0))
        roof["peak"] = hbm
        roof["peak_source"] = "MEASURED_PEAKS.json hbm_gbs (measured copy)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        roof["frac"] = roof["achieved"] / hbm
        try:     # measured DRAM bytes per launch of the same kernel at the same size, from the committed ncu --set full capture
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["adam_clamp_flat_kernel"]
            if tr["elements"] == roof["elements"]:
                roof["traffic"] = tr["dram_bytes_read"] + tr["dram_bytes_write"]
                roof["traffic_source"] = tr["source"]
        except Exception:
            pass
        for o in roof["other_kernels"]:
            if "achieved" in o:
                o["frac"] = o["achieved"] / hbm
                o["fwd"]["frac"] = o["fwd"]["achieved"] / hbm
        out = {
            "metric": METRIC, "value": world * B * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": dict(train_config(args, world), loss_from_lowres=bool(ts.loss_from_lowres)),
            "e2e": {"value": world * B * args.steps / (ms_e2e * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": imgs_h[0].numel() * 4 + gts_h[0].numel() * 4, "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
        }
        if not args.no_cpu_baseline and world == 1:      # the CPU baseline is reported at N = 1 only
            ips, per = time_cpu(args.cpu_batch, S, 3, 1)
            out["cpu_baseline"] = {"value": ips, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                   "sample": f"3 timed steps (1 warm-up) of batch {args.cpu_batch} @ {S}^2: stock Res2Net-50 on torch-CPU + oracle head/loss, fp32"}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_infer(args, P, model, dev, world, rank, local):
    """BASELINE.json's "infer images/sec @352^2": batched test-time path (binary_seg/MyTest_med.py:28-42) -- forward in eval mode +
    fused uint8 tail -- replicas only across GPUs (no collective).  Same timing rules as the training arm."""
    import torch.distributed as dist
    from pranet_v2_b200.train import InferStep
    B, S = args.batch, args.size
    inf = InferStep(model, device=dev, autocast=(args.precision == "bf16"), use_graph=not args.no_graph)
    g = torch.Generator().manual_seed(2000 + rank)
    nbuf = 4
    imgs_h = [torch.randn(B, 3, S, S, generator=g).pin_memory() for _ in range(nbuf)]
    imgs_d = [t.to(dev) for t in imgs_h]
    out_h = torch.empty(B, S, S, dtype=torch.uint8).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        inf.predict_device(imgs_d[i % nbuf])
    barrier()
    if rank == 0:
        sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        inf.predict_device(imgs_d[i % nbuf])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    for i in range(2):
        inf.predict_host(imgs_h[i % nbuf], out_h)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for i in range(args.steps):
        inf.predict_host(imgs_h[i % nbuf], out_h)
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()
    if rank == 0:
        print(json.dumps({
            "metric": "infer images/sec @352^2 (PraNet-V2 Res2Net-50 forward + fused uint8 tail)", "value": world * B * args.steps / (ms * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": f"PraNet-V2 Res2Net-50 batched inference (eval forward + p2+p3+p4+p5 / resize / sigmoid / min-max / uint8 tail), per-GPU batch {B} @ {S}^2",
                       "global_batch": world * B, "parallelism": f"replicas x{world}", "cuda_graph": not args.no_graph, "l2": "inputs rotate over 4 batches"},
            "e2e": {"value": world * B * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": imgs_h[0].numel() * 4, "d2h_bytes_per_step": out_h.numel()},
            "gpu_launches": int(inf.pv2_launches_per_step * args.steps), "clocks": clocks}))
    if world > 1:
        dist.destroy_process_group()


def roofline_adam(P, dev, n, iters=20):
    """pv2_adam_clamp_flat over the step's flat parameter layout: the largest-traffic pv2 launch of the training step
    (28 algorithmic bytes per parameter: p, g, m, v read; p, m, v written).  Timed with CUDA events around back-to-back C-ABI
    launches on torch's current stream; the four buffers (~130 MB each at 32.5 M parameters) exceed the 126 MB L2 several
    times over, so every launch streams from HBM."""
    lib = P._lib.load()
    st = torch.cuda.current_stream().cuda_stream
    n = (int(n) + 3) // 4 * 4
    p = torch.randn(n, device=dev)
    g = torch.randn(n, device=dev)
    m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    step = torch.zeros(1, dtype=torch.int64, device=dev)
    ticket = torch.zeros(1, dtype=torch.int32, device=dev)

    def run():
        P._lib.check(lib.pv2_adam_clamp_flat(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, step.data_ptr(), ticket.data_ptr(),
                                             1e-4, 0.9, 0.999, 1e-8, 0.0, 0, 0.5, 1.0, st), "adam_clamp_flat")
    for _ in range(3):
        run()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        run()
    b.record()
    torch.cuda.synchronize()
    t = a.elapsed_time(b) * 1e-3 / iters
    return {"kernel": "adam_clamp_flat_kernel (clip_gradient + Adam, flat parameters)", "bound": "hbm", "achieved": 28.0 * n / t / 1e9, "unit": "GB/s",
            "bytes_per_launch": 28 * n, "elements": n, "avg_ms": t * 1e3, "traffic": None,
            "note": f"{iters} back-to-back C-ABI launches between two CUDA events; 4 x {4 * n / 1e6:.0f} MB buffers (> L2)"}


def roofline_structure_loss(P, dev, B, S, iters=24):
    """structure_loss x4 backward: the largest-traffic pv2 launch of the step, timed live with CUDA events
    around back-to-back C-ABI launches on torch's current stream (inputs rotate over sets larger than L2).
    Algorithmic bytes / launch = P * (4 [mask] + 4 scales * (8 read + 8 written)) fp32, P = B*S*S."""
    from pranet_v2_b200 import synthetic as synth
    lib = P._lib.load()
    st = torch.cuda.current_stream().cuda_stream
    shape = (B, 1, S, S)
    m = synth.ellipse_masks(B, S, S, 3).to(dev)
    nset = 6    # 6 sets x (8 logits + 8 grads) x 7.9 MB  ~ 760 MB at B=16: far beyond the 126 MB L2
    logits = [[torch.randn(shape, device=dev) for _ in range(8)] for _ in range(nset)]
    grads = [[torch.empty(shape, device=dev) for _ in range(8)] for _ in range(nset)]
    ws_bytes = lib.pv2_structure_loss_workspace_bytes(B, S, S, 4)
    ws = torch.empty(ws_bytes // 4, device=dev)
    loss = torch.empty(4, device=dev)
    gl = torch.ones(4, device=dev)
    packs = []
    for j in range(nset):
        packs.append((P._lib.ptr_array(logits[j][:4]), P._lib.ptr_array(logits[j][4:]), P._lib.ptr_array(grads[j][:4]), P._lib.ptr_array(grads[j][4:])))

    def fwd(j):
        (pp, _), (pb, _), _, _ = packs[j]
        P._lib.check(lib.pv2_structure_loss_fwd(pp, pb, m.data_ptr(), None, 4, B, S, S, 0, loss.data_ptr(), ws.data_ptr(), ws_bytes, st), "fwd")

    def bwd(j):
        (pp, _), (pb, _), (dp, _), (dq, _) = packs[j]
        P._lib.check(lib.pv2_structure_loss_bwd(pp, pb, m.data_ptr(), None, gl.data_ptr(), dp, dq, 4, B, S, S, 0, ws.data_ptr(), ws_bytes, st), "bwd")

    def timed(fn):
        for j in range(nset):
            fn(j)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for it in range(iters):
            fn(it % nset)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) * 1e-3 / iters

    fwd(0)
    t_f, t_b = timed(fwd), timed(bwd)
    px = B * S * S
    bytes_bwd, bytes_fwd = px * (4 + 4 * 16), px * (4 + 4 * 8)
    return {"kernel": "structure_loss_bwd_kernel<float> (x4 scales)", "bound": "hbm", "achieved": bytes_bwd / t_b / 1e9, "unit": "GB/s",
            "bytes_per_launch": bytes_bwd, "avg_ms": t_b * 1e3, "traffic": None,
            "fwd": {"kernel": "structure_loss_fwd_kernel<float> (x4 scales) + finalize", "achieved": bytes_fwd / t_f / 1e9,
                    "bytes_per_launch": bytes_fwd, "avg_ms": t_f * 1e3},
            "note": f"{iters} back-to-back C-ABI launches between two CUDA events; inputs rotate over {nset} sets (> L2)"}


def roofline_bilinear(P, dev, B, S, n=24, reps=5):
    """The eight final upsamples of a step (pranet.py:349-350,370-371,392-393,414-415: x8, x16, x32, x8 for fg and bg) as the head
    issues them -- one pv2_bilinear_multi_fwd launch, one pv2_bilinear_multi_bwd launch.  These launches take 12-27 us, less than a
    ctypes call costs on the host, so `n` of them are captured in a CUDA graph (maps rotate over 4 sets, 4 x 63 MB > L2) and the
    replay is timed with CUDA events.  Algorithmic bytes / launch = (8 full-resolution maps + their low-res sources) * 4."""
    import ctypes
    from pranet_v2_b200.ops import PV2_F32, _ratio
    lib = P._lib.load()
    scs = (8, 16, 32, 8, 8, 16, 32, 8)
    nset = 4
    lows = [[torch.randn(B, 1, S // s, S // s, device=dev) for s in scs] for _ in range(nset)]
    his = [[torch.randn(B, 1, S, S, device=dev) for _ in scs] for _ in range(nset)]
    ihs = (ctypes.c_int * 8)(*[S // s for s in scs])
    rr = (ctypes.c_float * 8)(*[_ratio(S // s, S, False, float(s)) for s in scs])
    pk = [(P._lib.ptr_array(lows[j]), P._lib.ptr_array(his[j])) for j in range(nset)]

    def fwd(j):
        (pl, _), (ph, _) = pk[j]
        P._lib.check(lib.pv2_bilinear_multi_fwd(pl, ph, ihs, ihs, rr, rr, 8, B, S, S, 0, PV2_F32, torch.cuda.current_stream().cuda_stream), "bilinear_multi_fwd")

    def bwd(j):
        (pl, _), (ph, _) = pk[j]
        P._lib.check(lib.pv2_bilinear_multi_bwd(ph, pl, ihs, ihs, rr, rr, 8, B, S, S, 0, PV2_F32, torch.cuda.current_stream().cuda_stream), "bilinear_multi_bwd")

    def timed(fn):
        fn(0)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for it in range(n):
                fn(it % nset)
        g.replay()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) * 1e-3 / (reps * n)

    t_f, t_b = timed(fwd), timed(bwd)
    nbytes = (8 * B * S * S + sum(B * (S // s) ** 2 for s in scs)) * 4
    return {"kernel": "bilinear_bwd_kernel<float> (8 final maps, one launch)", "bound": "hbm", "achieved": nbytes / t_b / 1e9, "unit": "GB/s",
            "bytes_per_launch": nbytes, "avg_ms": t_b * 1e3, "traffic": None,
            "fwd": {"kernel": "bilinear_fwd_kernel<float> (8 final maps, one launch)", "achieved": nbytes / t_f / 1e9,
                    "bytes_per_launch": nbytes, "avg_ms": t_f * 1e3},
            "note": f"{n} launches captured in a CUDA graph, {reps} replays between two CUDA events; maps rotate over {nset} sets (> L2)"}


if __name__ == "__main__":
    main()
