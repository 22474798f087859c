"""Optimizer tail: clip_gradient + Adam / AdamW as one flat stream.
CPU: the numpy oracle is pinned against the installed torch.optim run on CPU; the flat layout keeps values, shapes, strides.
GPU: pv2_adam_clamp_flat through the C ABI == oracle == torch.optim on the device.  Tolerance: 2e-6 absolute on parameters
of O(1) after 4 steps at lr 1e-2 (fp32 rounding of differently ordered but algebraically equal expressions)."""
import numpy as np
import pytest
import torch

from oracle import optim_oracle as OO

SHAPES = [(7,), (5, 3, 3, 3), (64, 32, 1, 1), (33,), (2, 9, 5, 5), (1,), (130, 17)]


def _tensors(seed):
    g = torch.Generator().manual_seed(seed)
    ps = [torch.randn(s, generator=g) for s in SHAPES]
    ps[1] = ps[1].contiguous(memory_format=torch.channels_last)          # the stock backbone's conv weights under channels_last
    ps[4] = ps[4].contiguous(memory_format=torch.channels_last)
    grads = [[torch.randn(s, generator=g) * (2.0 if i % 2 else 0.05) for s in SHAPES] for i in range(4)]   # some beyond the clamp
    return ps, grads


@pytest.mark.parametrize("decoupled,wd", [(False, 0.0), (True, 1e-4), (False, 1e-2)])
def test_oracle_matches_torch_optim_cpu(decoupled, wd):
    ps, grads = _tensors(0)
    tp = [torch.nn.Parameter(p.clone()) for p in ps]
    opt = (torch.optim.AdamW if decoupled else torch.optim.Adam)(tp, lr=1e-2, weight_decay=wd, foreach=False)
    op = [p.numpy().copy() for p in ps]
    om = [np.zeros_like(a) for a in op]
    ov = [np.zeros_like(a) for a in op]
    for it, gs in enumerate(grads, start=1):
        for p, g in zip(tp, gs):
            p.grad = g.clone()
            p.grad.data.clamp_(-0.5, 0.5)
        opt.step()
        for i, g in enumerate(gs):
            op[i], om[i], ov[i] = OO.clamp_adam_step(op[i], g.numpy(), om[i], ov[i], it, lr=1e-2, weight_decay=wd, decoupled=decoupled, clip=0.5)
    for a, b in zip(tp, op):
        np.testing.assert_allclose(a.detach().numpy(), b, rtol=0, atol=2e-6)


def test_flat_layout_cpu():
    from pranet_v2_b200.train import FlatParams
    ps, _ = _tensors(1)
    params = [torch.nn.Parameter(p.clone()) for p in ps]
    fp = FlatParams(params, "cpu")
    assert fp.n % 4 == 0 and all(o % 4 == 0 for o in fp.offsets)
    for p, ref, gv, off in zip(params, ps, fp.views, fp.offsets):
        assert torch.equal(p.detach(), ref) and p.stride() == ref.stride() and gv.stride() == ref.stride()
        assert p.data_ptr() == fp.p.data_ptr() + 4 * off and gv.data_ptr() == fp.g.data_ptr() + 4 * off
    fp.gather([torch.full_like(p, float(i + 1)) for i, p in enumerate(params)])
    for i, (p, off) in enumerate(zip(params, fp.offsets)):
        seg = fp.g[off:off + p.numel()]
        assert torch.all(seg == i + 1)
    pad = fp.g.clone()
    for p, off in zip(params, fp.offsets):
        pad[off:off + p.numel()] = 0
    assert torch.all(pad == 0), "padding elements must carry zero gradients"


@pytest.mark.gpu
@pytest.mark.parametrize("decoupled,wd,world", [(False, 0.0, 1), (True, 1e-4, 1), (False, 1e-2, 2)])
def test_adam_clamp_flat_gpu(decoupled, wd, world):
    from pranet_v2_b200.train import FlatParams
    ps, grads = _tensors(2)
    params = [torch.nn.Parameter(p.clone().cuda()) for p in ps]
    fp = FlatParams(params, "cuda", lr=1e-2, weight_decay=wd, decoupled=decoupled, clip=0.5)
    fp.world = world                                                      # grad_scale = 1/world (the all-reduce itself is NCCL's)
    tp = [torch.nn.Parameter(p.clone().cuda()) for p in ps]
    opt = (torch.optim.AdamW if decoupled else torch.optim.Adam)(tp, lr=1e-2, weight_decay=wd)
    op = [p.numpy().copy() for p in ps]
    om = [np.zeros_like(a) for a in op]
    ov = [np.zeros_like(a) for a in op]
    for it, gs in enumerate(grads, start=1):
        fp.gather([g.cuda() * world for g in gs])                         # the summed message of `world` identical ranks
        fp.update()
        for p, g in zip(tp, gs):
            p.grad = g.cuda().clamp_(-0.5, 0.5)
        opt.step()
        for i, g in enumerate(gs):
            op[i], om[i], ov[i] = OO.clamp_adam_step(op[i], g.numpy() * world, om[i], ov[i], it, lr=1e-2, weight_decay=wd,
                                                     decoupled=decoupled, clip=0.5, grad_scale=1.0 / world)
    torch.cuda.synchronize()
    assert int(fp.step.item()) == len(grads) and int(fp.ticket.item()) == 0
    for p, t, o in zip(params, tp, op):
        np.testing.assert_allclose(p.detach().cpu().numpy(), o, rtol=0, atol=2e-6)
        np.testing.assert_allclose(p.detach().cpu().numpy(), t.detach().cpu().numpy(), rtol=0, atol=2e-6)


@pytest.mark.gpu
def test_adam_clamp_flat_large_and_graph_replay():
    """32.5 M elements (the PraNet-V2 Res2Net-50 parameter count): every element updated exactly once per launch, and the
    device-side step counter advances under CUDA-graph replay."""
    from pranet_v2_b200 import _lib
    lib = _lib.load()
    n = 32_550_004
    g = torch.Generator(device="cuda").manual_seed(0)
    p = torch.randn(n, device="cuda", generator=g)
    gr = torch.randn(n, device="cuda", generator=g)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    step = torch.zeros(1, dtype=torch.int64, device="cuda")
    ticket = torch.zeros(1, dtype=torch.int32, device="cuda")
    ref_p = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([ref_p], lr=1e-3)

    def launch():
        _lib.check(lib.pv2_adam_clamp_flat(p.data_ptr(), gr.data_ptr(), m.data_ptr(), v.data_ptr(), n, step.data_ptr(), ticket.data_ptr(),
                                           1e-3, 0.9, 0.999, 1e-8, 0.0, 0, 0.5, 1.0, torch.cuda.current_stream().cuda_stream), "adam")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        launch()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            launch()
    torch.cuda.synchronize()
    graph.replay()
    graph.replay()
    torch.cuda.synchronize()
    assert int(step.item()) == 3                                          # 1 eager launch + 2 replays (capture itself does not execute)
    for _ in range(3):
        ref_p.grad = gr.clamp(-0.5, 0.5)
        opt.step()
    assert (p - ref_p.detach()).abs().max().item() <= 2e-6
