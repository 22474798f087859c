"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: batch sharding gives the global mean loss, and the
flat-bucket all-reduce + clamp + Adam of train.TrainStep keeps replicas identical.  The kernels themselves are
CUDA-only, so the per-rank 'model' here is a tiny CPU module; what is under test is the data-parallel plumbing."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pranet_v2_b200.train import FlatGradBucket
    torch.manual_seed(0)                               # identical replicas
    model = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(4, 1, 1))
    params = list(model.parameters())
    bucket = FlatGradBucket(params, device="cpu")
    opt = torch.optim.Adam(params, 1e-2)
    g = torch.Generator().manual_seed(100)
    x_all, y_all = torch.randn(8, 3, 8, 8, generator=g), torch.randn(8, 1, 8, 8, generator=g)
    shard = slice(rank * 4, rank * 4 + 4)              # batch sharding, equal shards
    for _ in range(3):
        for p in params:
            p.grad = None
        loss = ((model(x_all[shard]) - y_all[shard]) ** 2).mean()
        loss.backward()
        bucket.gather([p.grad for p in params])
        bucket.all_reduce_mean()                       # the path's only collective
        bucket.clamp_(0.5)
        bucket.install(params)
        opt.step()
    # reference: single process on the full batch
    torch.manual_seed(0)
    ref = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(4, 1, 1))
    ropt = torch.optim.Adam(ref.parameters(), 1e-2)
    for _ in range(3):
        ropt.zero_grad()
        ((ref(x_all) - y_all) ** 2).mean().backward()
        for p in ref.parameters():
            p.grad.clamp_(-0.5, 0.5)
        ropt.step()
    err = max((a - b).abs().max().item() for a, b in zip(model.parameters(), ref.parameters()))
    flat = torch.cat([p.detach().flatten() for p in params])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    q.put((rank, err, max((gathered[0] - t).abs().max().item() for t in gathered)))
    dist.destroy_process_group()


def test_flat_bucket_allreduce_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, err_vs_single, spread in res:
        assert spread == 0.0, "replicas diverged"
        assert err_vs_single < 1e-6, f"2-rank data parallel != single process on the full batch ({err_vs_single})"


def _worker_flat(rank, world, port, q):
    """The product's multi-GPU data path on 2 gloo ranks: FlatParams layout -> gather -> ONE all-reduce (sum) of the flat
    gradient buffer -> the optimizer tail with grad_scale = 1/world.  The tail kernel itself is CUDA-only, so its CPU oracle
    stands in for it here; what is under test is the sharding, the message layout and the 1/world placement."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import optim_oracle as OO
    from pranet_v2_b200.train import FlatParams
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Conv2d(3, 5, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(5, 1, 1))
    model = model.to(memory_format=torch.channels_last)
    params = list(model.parameters())
    fp = FlatParams(params, "cpu", lr=1e-2, clip=0.5)
    assert fp.world == world
    g = torch.Generator().manual_seed(100)
    x_all, y_all = torch.randn(8, 3, 8, 8, generator=g), torch.randn(8, 1, 8, 8, generator=g)
    shard = slice(rank * 4, rank * 4 + 4)
    for it in range(1, 4):
        for p in params:
            p.grad = None
        ((model(x_all[shard]) - y_all[shard]) ** 2).mean().backward()
        fp.gather([p.grad for p in params])
        fp.all_reduce_sum()
        newp, newm, newv = OO.clamp_adam_step(fp.p.numpy(), fp.g.numpy(), fp.m.numpy(), fp.v.numpy(), it, lr=fp.lr, clip=fp.clip,
                                              grad_scale=1.0 / fp.world)
        fp.p.copy_(torch.from_numpy(newp)); fp.m.copy_(torch.from_numpy(newm)); fp.v.copy_(torch.from_numpy(newv))
    torch.manual_seed(0)
    ref = torch.nn.Sequential(torch.nn.Conv2d(3, 5, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(5, 1, 1))
    ropt = torch.optim.Adam(ref.parameters(), 1e-2)
    for _ in range(3):
        ropt.zero_grad()
        ((ref(x_all) - y_all) ** 2).mean().backward()
        for p in ref.parameters():
            p.grad.clamp_(-0.5, 0.5)
        ropt.step()
    err = max((a - b).abs().max().item() for a, b in zip(model.parameters(), ref.parameters()))
    gathered = [torch.empty_like(fp.p) for _ in range(world)]
    dist.all_gather(gathered, fp.p)
    q.put((rank, err, max((gathered[0] - t).abs().max().item() for t in gathered)))
    dist.destroy_process_group()


def test_flat_params_allreduce_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_flat, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, err_vs_single, spread in res:
        assert spread == 0.0, "replicas diverged"
        assert err_vs_single < 2e-6, f"2-rank flat-parameter data parallel != single process on the full batch ({err_vs_single})"
