"""Is a replay of the head-step graph bound by the host's enqueue time or by the GPU?  Prints the host time of graph.replay()
(no sync) next to the device time per replay, for the multi-branch graph and the single-stream one (PV2_STREAMS=0)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pranet_v2_b200 as P
from pranet_v2_b200 import synthetic

B, S = 16, 352
dev = "cuda:0"
P.set_precision("bf16")
torch.manual_seed(0)
model = P.PraNet_V2(num_class=1).to(dev).train()
g = torch.Generator(device="cpu").manual_seed(1)
feats = [torch.relu(torch.randn(B, c, S // s, S // s, generator=g)).to(dev).bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
         for c, s in zip((512, 1024, 2048), (8, 16, 32))]
gt = synthetic.ellipse_masks(B, S, S, 3).to(dev)
params = model.head_parameters()
prep_stream = torch.cuda.Stream()

def step():
    for p in params:
        p.grad = None
    for f in feats:
        f.grad = None
    prepared = P.ops.structure_loss_prepare(gt, prep_stream)
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
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    step()
for _ in range(5):
    graph.replay()
torch.cuda.synchronize()
N = 30
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
a.record()
for _ in range(N):
    graph.replay()
b.record()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"streams={os.environ.get('PV2_STREAMS', '1')}: host enqueue {1e3 * (t1 - t0) / N:.3f} ms/replay, device {a.elapsed_time(b) / N:.3f} ms/replay, wall incl. sync {1e3 * (t2 - t0) / N:.3f} ms/replay")
# one replay at a time, synchronised: the latency of a single step
ts = []
for _ in range(10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    graph.replay()
    torch.cuda.synchronize()
    ts.append(time.perf_counter() - t0)
print(f"single synchronised replay: median {1e3 * sorted(ts)[len(ts) // 2]:.3f} ms")
