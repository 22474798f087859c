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
# BASELINE.json configs: [1] Res2Net-50 training (the headline), [2] PVTv2-b2 training / inference, [3] EMCAD + DSRA multiclass
CONFIGS = {
    "res2net": dict(model="PraNet_V2", kw=dict(num_class=1), size=352, task="binary", label="PraNet-V2 Res2Net-50"),
    "pvt": dict(model="PVT_PraNet_V2", kw=dict(num_class=1), size=352, task="binary", label="PVT-PraNet-V2 (PVTv2-b2)"),
    "emcad": dict(model="EMCADNet", kw=dict(num_classes=9, encoder="pvt_v2_b2", pretrain=False, dual=True), size=224, task="multiclass",
                  label="EMCAD (PVTv2-b2) + DSRA, 9 classes"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="res2net", choices=list(CONFIGS), help="res2net (headline, BASELINE config 2) | pvt (config 3) | emcad (config 4)")
    ap.add_argument("--batch", type=int, default=16, help="per-GPU batch")
    ap.add_argument("--size", type=int, default=None, help="input size (default: 352, emcad: 224)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary records (inference, PVT, EMCAD) embedded in the N=1 line")
    ap.add_argument("--no-eager-reference", action="store_true", help="skip the reference-in-eager-PyTorch-on-this-GPU leg")
    ap.add_argument("--ref-seconds", type=float, default=150.0, help="wall-clock budget of the reference arm's timed steps (a bounded sample)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--mode", default="train", choices=["train", "infer"],
                    help="train (default, the headline metric) or infer: batched test-time path, images/s of uint8 saliency maps")
    ap.add_argument("--profile-only", action="store_true",
                    help="run the warm-up + timed steps only (no e2e / roofline / CPU legs) and exit: the command ncu wraps for the launch list")
    a = ap.parse_args()
    if a.size is None:
        a.size = CONFIGS[a.config]["size"]
    return a


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
# Reference arm / CPU baseline: the reference's OWN modules (oracle/_ref, staged byte for byte by build()) on the host cores;
# the oracle port when they are not there.  fp32, all host threads, the benchmarked batch.
# ------------------------------------------------------------------------------------------------
def _labels(B, S, C, seed):
    """9-class blob maps: background 0 ~ 70 %, the other classes as random ellipses."""
    import numpy as np
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:S, 0:S]
    out = np.zeros((B, S, S), dtype=np.int64)
    for b in range(B):
        for c in range(1, C):
            cy, cx = rng.uniform(0.15, 0.85, 2) * S
            ry, rx = rng.uniform(0.04, 0.14, 2) * S
            out[b][((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0] = c
    return torch.from_numpy(out)


def reference_step_factory(config: str, batch: int, size: int, device="cpu", mode="train", variant="fp32"):
    """One step of the reference's path with the reference's own modules when they are staged (kind 'reference'), else the oracle
    port (kind 'port').  train: model forward, the loss block (MyTrain_med.py:76-82 / EMCAD trainer.py:105-140), backward,
    clip_gradient (binary), Adam / AdamW.  infer: eval forward + the test-time rule.  Returns (step, kind, description)."""
    from oracle import ref_import as R
    from oracle import synth
    cfg = CONFIGS[config]
    dev = torch.device(device)
    if dev.type == "cpu":
        torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(1000)
    if R.available():
        kind = "reference"
        if config == "emcad":
            model = R.emcad_networks().EMCADNet(**cfg["kw"])
            x = torch.randn(batch, 1, size, size, generator=g)
            y = _labels(batch, size, 9, 1000)
        else:
            model = R.build_binary(cfg["model"], **cfg["kw"])
            x = torch.randn(batch, 3, size, size, generator=g)
            y = synth.ellipse_masks(batch, size, size, 1000)
        model = model.to(dev)
        x, y = x.to(dev), y.to(dev)
        if variant == "bf16_cl":
            model = model.to(memory_format=torch.channels_last)
            x = x.contiguous(memory_format=torch.channels_last)
        autocast = variant == "bf16_cl"
        if mode == "infer":
            model.eval()

            def step():
                with torch.no_grad(), torch.autocast(dev.type, dtype=torch.bfloat16, enabled=autocast):
                    o = model(x)
                    res = (o[0] + o[1] + o[2] + o[3]).float().sigmoid()          # MyTest_med.py:35-39
                    res = (res - res.amin((1, 2, 3), keepdim=True)) / (res.amax((1, 2, 3), keepdim=True) - res.amin((1, 2, 3), keepdim=True) + 1e-8)
                    return (res * 255).to(torch.uint8)
            return step, kind, f"{cfg['label']} eval forward + test-time rule, the reference's own modules"
        model.train()
        if config == "emcad":
            powerset, DiceLoss, one_hot = R.emcad_loss_pieces()
            ce, dice, bce = torch.nn.CrossEntropyLoss(), DiceLoss(9), torch.nn.BCEWithLogitsLoss()
            opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-4)     # EMCAD/trainer.py:86
            subsets = [s for s in powerset(list(range(4))) if s]

            def step():
                opt.zero_grad(set_to_none=True)
                with torch.autocast(dev.type, dtype=torch.bfloat16, enabled=autocast):
                    P = model(x, mode="train")
                P = [p.float() for p in P]
                bg_mask = one_hot(y.cpu(), 9).to(dev).float()                           # trainer.py:22-29,99-103 (built on the CPU there)
                loss = 0.0
                for sub in subsets:                                                     # trainer.py:123-140
                    iout = sum(P[i] for i in sub)
                    ibg = sum(P[i + 4] for i in sub)
                    loss = loss + 0.5 * ce(iout, y) + 0.7 * dice(iout, y, softmax=True) + 0.3 * bce(ibg, bg_mask)
                loss.backward()
                opt.step()
                return float(loss)
            return step, kind, "EMCADNet(dual) train step: forward, 15-subset dual loss, backward, AdamW -- the reference's own modules"
        sl = R.structure_loss()
        opt = torch.optim.Adam(model.parameters(), 1e-4)                                # MyTrain_med.py:148-149

        def step():
            opt.zero_grad(set_to_none=True)
            with torch.autocast(dev.type, dtype=torch.bfloat16, enabled=autocast):
                o = model(x)
            o = [t.float() for t in o]
            loss = sum(sl(o[i], o[i + 4], y, 1 - y) for i in range(4))                  # MyTrain_med.py:74,78-82
            loss.backward()
            for grp in opt.param_groups:                                                # utils/utils.py:7-17 clip_gradient(optimizer, 0.5)
                for p in grp["params"]:
                    if p.grad is not None:
                        p.grad.data.clamp_(-0.5, 0.5)
            opt.step()
            return float(loss)
        return step, kind, f"{cfg['label']} train step (forward, 4x structure_loss, backward, clip_gradient, Adam), the reference's own modules"
    # ---- oracle port (the staged reference is missing) ----
    if config != "res2net" or mode != "train" or dev.type != "cpu":
        raise RuntimeError("oracle/_ref is not staged (run __graft_entry__.build() in the builder container): only the Res2Net training "
                           "port is available")
    from oracle import dsra_oracle as O
    from oracle import templates
    from pranet_v2_b200.backbones import Res2Net50
    bb = Res2Net50().train()
    sd = synth.synth_state_dict(templates.pranet_head(num_class=1), seed=0)
    params = [v.requires_grad_(True) for k, v in sd.items() if v.dtype.is_floating_point and "running" not in k]
    opt = torch.optim.Adam(list(bb.parameters()) + params, 1e-4)
    x = torch.randn(batch, 3, size, size, generator=g)
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
    return step, "port", "oracle port: stock Res2Net-50 on torch-CPU + oracle head / loss (the staged reference modules are missing)"


def time_cpu(config, batch, size, steps, warmup, budget_s, mode="train"):
    """(images/s, s/step, steps timed, kind, description): a bounded sample -- at most `steps` steps, at least one, stopping once
    `budget_s` seconds of timed work are spent (a batch-16 Res2Net training step is seconds of CPU time)."""
    step, kind, desc = reference_step_factory(config, batch, size, "cpu", mode)
    t_w = time.perf_counter()
    for _ in range(max(0, warmup)):
        step()
        if time.perf_counter() - t_w > budget_s / 3:
            break
    done, t0 = 0, time.perf_counter()
    while done < steps:
        step()
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return batch * done / dt, dt / done, done, kind, desc


def workload(args, world):
    """The workload both arms are quoted on."""
    cfg = CONFIGS[args.config]
    if args.mode == "infer":
        what = f"{cfg['label']} batched inference (eval forward + test-time rule -> uint8 maps)"
    elif cfg["task"] == "multiclass":
        what = f"{cfg['label']} train step (fwd + 15-subset dual loss + bwd + AdamW)"
    else:
        what = f"{cfg['label']} train step (fwd + 4x structure_loss + bwd + clamp + Adam)"
    return {"workload": f"{what}, per-GPU batch {args.batch} @ {args.size}^2, random init, synthetic data",
            "global_batch": world * args.batch, "parallelism": f"dp{world}" if args.mode == "train" else f"replicas x{world}"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ips, per, done, kind, desc = time_cpu(args.config, args.batch, args.size, args.steps, min(args.warmup, 1), args.ref_seconds, args.mode)
    cores = torch.get_num_threads()
    sample = f"{done} timed step(s) of batch {args.batch} @ {args.size}^2 (<= {args.ref_seconds:.0f} s budget; {desc}; fp32, {cores} host threads)"
    cfg = dict(workload(args, int(os.environ.get("WORLD_SIZE", "1"))), device="host CPU", cuda_graph=False)
    print(json.dumps({
        "impl": "reference", "metric": METRIC if (args.config == "res2net" and args.mode == "train") else cfg["workload"], "value": ips, "unit": UNIT,
        "n_gpus": args.gpus, "steps": done, "warmup": min(args.warmup, 1), "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": ips, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def gpu_eager_reference(args, dev):
    """The like-for-like "before" number: the reference's own modules in eager PyTorch on THIS GPU, same batch --
    fp32 with cuDNN off (as the reference ships: MyTrain_med.py:16) and bf16 autocast + channels_last with cuDNN on, the latter
    also replayed from a CUDA graph.  Full-step images/s, and the head + loss share as (full fwd+bwd) - (backbone-only fwd+bwd)."""
    from oracle import ref_import as R
    if not R.available():
        return {"unavailable": "oracle/_ref not staged (run __graft_entry__.build() where /root/reference is mounted)"}
    out = {"batch": args.batch, "size": args.size, "note": "reference modules (oracle/_ref), eager PyTorch on this GPU; CUDA events, 3 warm-up + 8 timed"}

    def timed(fn, n=8, warm=3):
        for _ in range(warm):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    saved = torch.backends.cudnn.enabled
    try:
        for variant, cudnn_on in (("fp32_cudnn_off", False), ("bf16_cl", True)):
            torch.backends.cudnn.enabled = cudnn_on
            try:
                step, _, _ = reference_step_factory(args.config, args.batch, args.size, dev, "train", "bf16_cl" if variant == "bf16_cl" else "fp32")
                ms = timed(step)
                name = "fp32_cudnn_off" if variant == "fp32_cudnn_off" else "bf16_autocast_channels_last_cudnn_on"
                out[name] = {"ms_per_step": ms, "images_per_s": args.batch / ms * 1e3}
            except Exception as exc:      # noqa: BLE001
                out[variant] = {"error": str(exc)[:200]}
            torch.cuda.empty_cache()
        # head + loss share on the bf16 / cuDNN-on variant: full fwd+bwd minus backbone-only fwd+bwd (the reference has no separable head)
        torch.backends.cudnn.enabled = True
        if args.config in ("res2net", "pvt"):
            try:
                from oracle import synth
                model = R.build_binary(CONFIGS[args.config]["model"], num_class=1).to(dev).to(memory_format=torch.channels_last).train()
                sl = R.structure_loss()
                x = torch.randn(args.batch, 3, args.size, args.size, device=dev).contiguous(memory_format=torch.channels_last)
                y = synth.ellipse_masks(args.batch, args.size, args.size, 1).to(dev)
                bb = model.backbone

                def full():
                    model.zero_grad(set_to_none=True)
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        o = model(x)
                    o = [t.float() for t in o]
                    sum(sl(o[i], o[i + 4], y, 1 - y) for i in range(4)).backward()

                def backbone_only():
                    model.zero_grad(set_to_none=True)
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        if args.config == "res2net":
                            t = bb.maxpool(bb.relu(bb.bn1(bb.conv1(x))))
                            x1 = bb.layer1(t); x2 = bb.layer2(x1); x3 = bb.layer3(x2); x4 = bb.layer4(x3)
                        else:
                            _, x2, x3, x4 = bb(x)
                    (x2.float().mean() + x3.float().mean() + x4.float().mean()).backward()
                t_full, t_bb = timed(full), timed(backbone_only)
                out["head_plus_loss_fwd_bwd_ms"] = {"full_fwd_bwd_ms": t_full, "backbone_only_fwd_bwd_ms": t_bb, "difference_ms": t_full - t_bb,
                                                    "variant": "bf16 autocast, channels_last, cuDNN on, eager"}
                # the same step replayed from a CUDA graph (launch overhead removed): what eager PyTorch can reach at best
                try:
                    side = torch.cuda.Stream()
                    side.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(side):
                        for _ in range(3):
                            full()
                    torch.cuda.current_stream().wait_stream(side)
                    torch.cuda.synchronize()
                    gr = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gr):
                        full()
                    t_g = timed(gr.replay)
                    with torch.cuda.graph(gr2 := torch.cuda.CUDAGraph()):
                        backbone_only()
                    t_gb = timed(gr2.replay)
                    out["head_plus_loss_fwd_bwd_ms"]["cuda_graph"] = {"full_fwd_bwd_ms": t_g, "backbone_only_fwd_bwd_ms": t_gb, "difference_ms": t_g - t_gb}
                except Exception as exc:      # noqa: BLE001
                    out["head_plus_loss_fwd_bwd_ms"]["cuda_graph"] = {"error": str(exc)[:200]}
                del model
            except Exception as exc:      # noqa: BLE001
                out["head_plus_loss_fwd_bwd_ms"] = {"error": str(exc)[:300]}
    finally:
        torch.backends.cudnn.enabled = saved
        torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------
def build_model(P, config):
    cfg = CONFIGS[config]
    return getattr(P, cfg["model"])(**cfg["kw"])


def make_batches(config, B, S, rank, nbuf=4):
    """Pinned host batches (inputs change step to step; together with the activations far beyond the 126 MB L2)."""
    from pranet_v2_b200 import synthetic as synth
    g = torch.Generator().manual_seed(1000 + rank)
    if CONFIGS[config]["task"] == "multiclass":
        imgs = [torch.randn(B, 1, S, S, generator=g).pin_memory() for _ in range(nbuf)]
        gts = [_labels(B, S, 9, 1000 + rank * 17 + i).pin_memory() for i in range(nbuf)]
    else:
        imgs = [torch.randn(B, 3, S, S, generator=g).pin_memory() for _ in range(nbuf)]
        gts = [synth.ellipse_masks(B, S, S, 1000 + rank * 17 + i).pin_memory() for i in range(nbuf)]
    return imgs, gts


def make_train_step(P, config, dev, precision, use_graph):
    from pranet_v2_b200.train import TrainStep
    cfg = CONFIGS[config]
    model = build_model(P, config)
    if cfg["task"] == "multiclass":     # EMCAD/trainer.py:86: AdamW(lr, weight_decay=1e-4), no gradient clipping
        return TrainStep(model, lr=1e-4, clip=0.0, autocast_backbone=(precision == "bf16"), device=dev, use_graph=use_graph,
                         weight_decay=1e-4, decoupled=True, task="multiclass", num_classes=9)
    return TrainStep(model, lr=1e-4, clip=0.5, autocast_backbone=(precision == "bf16"), device=dev, use_graph=use_graph)


def time_train(ts, imgs_d, gts_d, steps, warmup, barrier):
    nbuf = len(imgs_d)
    for i in range(warmup):
        ts.step_device(imgs_d[i % nbuf], gts_d[i % nbuf])
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        ts.step_device(imgs_d[i % nbuf], gts_d[i % nbuf])
    e1.record()
    barrier()
    return e0.elapsed_time(e1)


def secondary_records(P, dev, args):
    """BASELINE.json's other configurations, measured in the same run and embedded in the N = 1 line (device-resident, CUDA events,
    3 warm-up + 10 timed steps each): Res2Net inference, PVT training + inference (config 3), EMCAD + DSRA training (config 4)."""
    from pranet_v2_b200.train import InferStep
    out = {}

    def sync():
        torch.cuda.synchronize()

    def infer(config):
        B, S = args.batch, CONFIGS[config]["size"]
        inf = InferStep(build_model(P, config), device=dev, autocast=(args.precision == "bf16"), use_graph=not args.no_graph)
        g = torch.Generator().manual_seed(7)
        xs = [torch.randn(B, 3, S, S, generator=g).to(dev) for _ in range(4)]
        for i in range(3):
            inf.predict_device(xs[i % 4])
        sync()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(10):
            inf.predict_device(xs[i % 4])
        b.record()
        sync()
        ms = a.elapsed_time(b) / 10
        return {"metric": f"infer images/sec @{S}^2 ({CONFIGS[config]['label']} eval forward + fused uint8 tail)", "value": B / ms * 1e3, "unit": UNIT,
                "ms_per_step": ms, "batch": B, "gpu_launches_per_step": int(inf.pv2_launches_per_step), "dtype": args.precision}

    def train(config):
        B, S = args.batch, CONFIGS[config]["size"]
        ts = make_train_step(P, config, dev, args.precision, not args.no_graph)
        imgs, gts = make_batches(config, B, S, 0)
        ms = time_train(ts, [t.to(dev) for t in imgs], [t.to(dev) for t in gts], 10, 3, sync) / 10
        return {"metric": f"train images/sec @{S}^2 ({workload(argparse.Namespace(config=config, mode='train', batch=B, size=S), 1)['workload']})",
                "value": B / ms * 1e3, "unit": UNIT, "ms_per_step": ms, "batch": B, "gpu_launches_per_step": int(ts.pv2_launches_per_step), "dtype": args.precision}

    for name, fn, cfg in (("infer_res2net", infer, "res2net"), ("train_pvt", train, "pvt"), ("infer_pvt", infer, "pvt"), ("train_emcad", train, "emcad")):
        try:
            out[name] = fn(cfg)
        except Exception as exc:      # noqa: BLE001 -- a secondary record must not take the headline down with it
            out[name] = {"error": f"{type(exc).__name__}: {str(exc)[:300]}"}
        torch.cuda.empty_cache()
    return out


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    import pranet_v2_b200 as P

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
    if args.mode == "infer":
        return run_infer(args, P, build_model(P, args.config), dev, world, rank, local)
    ts = make_train_step(P, args.config, dev, args.precision, not args.no_graph)
    nbuf = 4
    imgs_h, gts_h = make_batches(args.config, B, S, rank, nbuf)
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

    if rank == 0:
        n_params = ts.bucket.n if hasattr(ts.bucket, "n") else ts.flat.numel()
        h2d = imgs_h[0].numel() * imgs_h[0].element_size() + gts_h[0].numel() * gts_h[0].element_size()
        cfg = dict(workload(args, world), cuda_graph=not args.no_graph, loss_from_lowres=bool(ts.loss_from_lowres),
                   l2="inputs rotate over 4 batches; per-step working set (activations + gradients) >> 126 MB L2")
        out = {
            "metric": METRIC if args.config == "res2net" else cfg["workload"], "value": world * B * args.steps / (ms * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic", "config": cfg,
            "e2e": {"value": world * B * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        del ts
        torch.cuda.empty_cache()
        # ---- roofline of the dominant hot-path kernel (the tcgen05 conv) and of the other pv2 kernels, timed live ----
        try:
            out["roofline"] = roofline_report(P, dev, B if args.config != "emcad" else 16, 352, n_params)
        except Exception as exc:      # noqa: BLE001
            out["roofline"] = {"error": f"{type(exc).__name__}: {str(exc)[:300]}"}
        if world == 1:
            if not args.no_cpu_baseline:      # the CPU baseline is reported at N = 1 only: a bounded sample of the same workload
                try:
                    ips, per, done, kind, desc = time_cpu(args.config, B, S, 2, 1, 25.0)
                    out["cpu_baseline"] = {"value": ips, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
                                           "sample": f"{done} timed step(s) (1 warm-up, <= 25 s) of batch {B} @ {S}^2: {desc}; fp32"}
                except Exception as exc:      # noqa: BLE001
                    out["cpu_baseline"] = {"error": str(exc)[:300]}
            if not args.no_eager_reference:
                out["gpu_eager_reference"] = gpu_eager_reference(args, dev)
            if not args.no_secondary and args.config == "res2net":
                out["secondary"] = secondary_records(P, dev, args)
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
        cfg = dict(workload(args, world), cuda_graph=not args.no_graph, l2="inputs rotate over 4 batches")
        print(json.dumps({
            "metric": f"infer images/sec @{S}^2 ({CONFIGS[args.config]['label']} forward + fused uint8 tail)", "value": world * B * args.steps / (ms * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.precision, "data": "synthetic", "config": cfg,
            "e2e": {"value": world * B * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": imgs_h[0].numel() * 4, "d2h_bytes_per_step": out_h.numel()},
            "gpu_launches": int(inf.pv2_launches_per_step * args.steps), "clocks": clocks}))
    if world > 1:
        dist.destroy_process_group()


def roofline_report(P, dev, B, S, n_params):
    """`roofline` of the bench line: the dominant hot-path kernel -- the persistent tcgen05 implicit-GEMM conv (102 of the ~440 pv2
    launches of a head step, the largest share of its kernel time) -- as the FLOP-weighted aggregate over the head's seven distinct
    forward shapes with the BatchNorm statistics fused, against the measured bf16 peak; `shapes` lists each; `other_kernels` carries
    the memory-bound kernels (structure loss fwd / bwd, final upsamples fwd / bwd, V1 reverse-attention scale, multiclass dual loss,
    optimizer tail) against the measured HBM copy peak.  Every number: CUDA events, launches captured in a CUDA graph (the launches
    take 5-40 us, less than a ctypes call costs on the host), inputs rotating over sets larger than the L2.  `traffic` = DRAM bytes
    per launch from the committed ncu --set full capture (profiles/ncu_traffic.json) where one exists."""
    import bench_head
    hbm, tfl, src = bench_head.peaks()
    rows = bench_head.kernel_table(P, dev, B, S, hbm, tfl)
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        traffic = {}
    convs = [r for r in rows if r["bound"] == "tensor" and r["kernel"].startswith("conv_fwd+bn_stats")]
    plain = {r["kernel"].replace("conv_fwd ", ""): r for r in rows if r["bound"] == "tensor" and r["kernel"].startswith("conv_fwd ")}
    fl, us = sum(r["flops"] for r in convs), sum(r["us"] for r in convs)
    roof = {"kernel": "conv_fwd2_kernel<bf16> (tcgen05 / TMEM / TMA implicit GEMM, BatchNorm statistics fused): FLOP-weighted over the head's 7 forward shapes",
            "bound": "tensor", "achieved": fl / (us * 1e-6) / 1e12, "peak": tfl, "unit": "TFLOP/s", "frac": fl / (us * 1e-6) / 1e12 / tfl,
            "peak_source": f"{src}: bf16_tflops (burst; each kernel is timed alone)", "flops_per_launch_set": fl, "us_per_launch_set": us,
            "traffic": traffic.get("conv_fwd2_kernel", {}).get("dram_bytes_per_launch_set"),
            "shapes": [{"shape": r["kernel"].replace("conv_fwd+bn_stats ", ""), "flops": r["flops"], "us": r["us"], "tflops": r["achieved_tflops"],
                        "frac": r["frac_of_bf16_peak"], "hbm_bytes": r.get("hbm_bytes"), "hbm_frac": r.get("frac_of_hbm_peak"), "roof": r.get("roof"),
                        "frac_of_own_roof": r.get("frac_of_own_roof"),
                        "us_without_bn_stats": plain.get(r["kernel"].replace("conv_fwd+bn_stats ", ""), {}).get("us")} for r in convs],
            "frac_of_own_roofs": sum(r.get("frac_of_own_roof", 0.0) * r["us"] for r in convs) / us,
            "note": "2*M*N*K un-padded; 24 launches per shape captured in a CUDA graph, inputs rotate over > 300 MB.  `frac` is against the bf16 tensor "
                    "peak for every shape; `roof` / `frac_of_own_roof` per shape use the lower of the tensor and the HBM roof (the 1x1 level GEMMs have "
                    "119-156 FLOP/B against a ridge of ~254), `frac_of_own_roofs` is their time-weighted mean"}
    others = []
    for r in rows:
        if r["bound"] != "hbm" or r["kernel"].startswith("yardstick"):
            continue
        key = r["kernel"].split(" (")[0]
        tr = traffic.get(key)
        others.append({"kernel": r["kernel"], "bound": "hbm", "achieved": r["achieved_gbs"], "peak": hbm, "unit": "GB/s", "frac": r["frac_of_hbm_peak"],
                       "bytes_per_launch": r["bytes"], "avg_ms": r["us"] * 1e-3,
                       "traffic": (tr["dram_bytes_read"] + tr["dram_bytes_write"]) if tr and "dram_bytes_read" in tr else None})
    for fn in (lambda: roofline_mc_loss(P, dev), lambda: roofline_adam(P, dev, n_params)):
        try:
            o = fn()
            o["peak"], o["frac"] = hbm, o["achieved"] / hbm
            tr = traffic.get(o.pop("traffic_key", ""), None)
            if tr and tr.get("elements", o.get("elements")) == o.get("elements"):
                o["traffic"] = tr["dram_bytes_read"] + tr["dram_bytes_write"]
            others.append(o)
        except Exception as exc:      # noqa: BLE001
            others.append({"error": str(exc)[:200]})
    roof["other_kernels"] = others
    roof["hbm_peak_source"] = f"{src}: hbm_gbs (measured copy)"
    return roof


def roofline_mc_loss(P, dev, B=16, C=9, S=224, n=12, reps=5):
    """Multiclass dual loss (EMCAD/trainer.py:123-140) at config 4's shape: forward and backward launches, algorithmic bytes
    fwd = P*8*4 (eight logit maps) + 8 B/px labels, bwd = the same reads + P*8*4 gradient bytes written."""
    g = torch.Generator().manual_seed(3)
    sets = [[torch.randn(B, C, S, S, generator=g).to(dev).requires_grad_(True) for _ in range(8)] for _ in range(3)]
    labels = _labels(B, S, C, 5).to(dev)

    def fwd(j):
        return P.mc_dual_loss(sets[j][:4], sets[j][4:], labels, C)

    def fb(j):
        for t in sets[j]:
            t.grad = None
        fwd(j).backward()

    def timed(fn):
        fn(0)
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn(0)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for it in range(n):
                fn(it % len(sets))
        gr.replay()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            gr.replay()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) * 1e-3 / (reps * n)

    with torch.no_grad():
        t_f = timed(lambda j: fwd(j))
    t_fb = timed(fb)
    px = B * C * S * S
    bytes_f = px * 8 * 4 + B * S * S * 8
    bytes_b = bytes_f + px * 8 * 4
    return {"kernel": f"mc_dual_loss fwd+bwd (15 subsets, {B}x{C}x{S}^2)", "bound": "hbm", "achieved": (bytes_f + bytes_b) / t_fb / 1e9, "unit": "GB/s",
            "bytes_per_launch": bytes_f + bytes_b, "avg_ms": t_fb * 1e3, "traffic": None, "traffic_key": "mc_dual_loss", "elements": px,
            "fwd": {"kernel": "mc_dual_loss fwd", "achieved": bytes_f / t_f / 1e9, "bytes_per_launch": bytes_f, "avg_ms": t_f * 1e3},
            "note": f"{n} fwd(+bwd) calls captured in a CUDA graph, {reps} replays; logits rotate over 3 sets"}


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
            "bytes_per_launch": 28 * n, "elements": n, "avg_ms": t * 1e3, "traffic": None, "traffic_key": "adam_clamp_flat_kernel",
            "note": f"{iters} back-to-back C-ABI launches between two CUDA events; 4 x {4 * n / 1e6:.0f} MB buffers (> L2)"}


if __name__ == "__main__":
    main()
