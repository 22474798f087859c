"""Inference tails from the low-res maps: binary_seg/MyTest_med.py:35-42 and EMCAD/utils/utils.py:285-296.

Golden vectors were produced by running the reference's own statements (oracle/make_golden.py group `tail`).
CPU: the oracle restatement reproduces them bit for bit.  GPU: the fused kernels, through the C ABI, agree with them.
Bars: uint8 outputs are integer work, but they sit behind a float sigmoid whose last-ulp rounding differs between libm
implementations (ATen CPU vs CUDA expf), so a value that lands within 1e-3 of an integer boundary may truncate to the
neighbour: every pixel must be within +-1, and >= 99.9 % of the pixels must be identical.  Label maps: identical wherever the
top-2 margin of the summed logits exceeds 1e-4, and >= 99.9 % identical overall (north_star mask agreement)."""
import numpy as np
import pytest
import torch

from oracle import dsra_oracle as O
from oracle import golden_cases as G


@pytest.mark.parametrize("name", list(G.TAIL_BINARY_CASES))
def test_oracle_binary_tail_matches_reference(name):
    maps, scales = G.tail_binary_inputs(name)
    got = O.infer_tail_binary(maps, scales, G.TAIL_BINARY_CASES[name]["gt"])
    assert np.array_equal(got, G.load(name)["out"])


@pytest.mark.parametrize("name", list(G.TAIL_MC_CASES))
def test_oracle_argmax_tail_matches_reference(name):
    fg, bg = G.tail_mc_inputs(name)
    got, _ = O.infer_tail_argmax(fg, bg, G.TAIL_MC_SCALES)
    assert np.array_equal(got, G.load(name)["labels"])


def _check_u8(got, want):
    d = np.abs(got.astype(np.int32) - want.astype(np.int32))
    assert d.max() <= 1, f"uint8 map off by {d.max()}"
    assert (d == 0).mean() >= 0.999, f"only {(d == 0).mean():.5f} of the pixels identical"


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(G.TAIL_BINARY_CASES))
def test_binary_tail_golden_gpu(name):
    import pranet_v2_b200 as P
    maps, scales = G.tail_binary_inputs(name)
    got = P.ops.infer_tail_binary([m.cuda() for m in maps], scales, G.TAIL_BINARY_CASES[name]["gt"]).cpu().numpy()
    _check_u8(got, G.load(name)["out"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(G.TAIL_MC_CASES))
def test_argmax_tail_golden_gpu(name):
    import pranet_v2_b200 as P
    fg, bg = G.tail_mc_inputs(name)
    got = P.ops.infer_tail_argmax([m.cuda() for m in fg], [m.cuda() for m in bg], G.TAIL_MC_SCALES).cpu().numpy()
    want, margin = O.infer_tail_argmax(fg, bg, G.TAIL_MC_SCALES)
    assert np.array_equal(want, G.load(name)["labels"])
    assert np.array_equal(got[margin > 1e-4], want[margin > 1e-4])
    assert (got == want).mean() >= 0.999


@pytest.mark.gpu
@pytest.mark.parametrize("B,S,gt", [(16, 352, None), (16, 352, (500, 574)), (3, 704, (704, 704)), (1, 256, (1, 7))])
def test_binary_tail_vs_oracle_full_size_gpu(B, S, gt):
    import pranet_v2_b200 as P
    from oracle import synth
    scales = list(G.TAIL_BIN_SCALES)
    maps = [synth.logits((B, 1, S // s, S // s), 11, f"m{k}", 2.0) for k, s in enumerate(scales)]
    got = P.ops.infer_tail_binary([m.cuda() for m in maps], scales, gt).cpu().numpy()
    want = O.infer_tail_binary(maps, scales, gt)
    _check_u8(got, want.reshape(got.shape))
    # size-independent properties: every image spans the full range (min-max normalisation), values are monotone in the logit
    assert all(got[b].min() == 0 and got[b].max() >= 254 for b in range(B) if got[b].size > 1)


@pytest.mark.gpu
def test_predict_uint8_matches_reference_rule_gpu():
    """model.predict_uint8 (forward + fused tail on the low-res maps) == the reference rule applied to the model's own
    full-resolution outputs."""
    import pranet_v2_b200 as P
    from oracle import synth
    torch.manual_seed(0)
    m = P.PraNet_V2(num_class=1).cuda().eval()
    feats = [f.cuda() for f in synth.backbone_features(2, 128, 3)]
    with torch.no_grad():
        low = m.forward_head(*feats, lowres=True)[:4]
        full = m.forward_head(*feats)[:4]
    got = P.ops.infer_tail_binary(low, m.final_scale_factors(), (150, 170)).cpu().numpy()
    out = (full[0] + full[1] + full[2] + full[3]).cpu()
    want = []
    for b in range(2):
        o = torch.nn.functional.interpolate(out[b:b + 1], size=(150, 170), mode="bilinear", align_corners=False)
        o = o.sigmoid().numpy().squeeze()
        o = (o - o.min()) / (o.max() - o.min() + 1e-8)
        want.append((o * 255).astype(np.uint8))
    _check_u8(got, np.stack(want))


@pytest.mark.gpu
def test_infer_step_graph_matches_eager_gpu():
    """InferStep (CUDA graph of eval forward + fused tail) reproduces the eager path and the reference rule on the model's own
    full-resolution outputs; a second batch replays the same graph."""
    import pranet_v2_b200 as P
    from pranet_v2_b200.train import InferStep
    from pranet_v2_b200 import synthetic
    torch.manual_seed(0)
    m = P.PraNet_V2(num_class=1)
    inf = InferStep(m, device="cuda:0", autocast=False, use_graph=True)
    for seed in (0, 1):
        x = synthetic.images(2, 128, seed).cuda()
        got = inf.predict_device(x, (100, 140)).cpu().numpy()
        with torch.no_grad():
            outs = inf.model(x.contiguous(memory_format=torch.channels_last))
        out = (outs[0] + outs[1] + outs[2] + outs[3]).float().cpu()
        want = []
        for b in range(2):
            o = torch.nn.functional.interpolate(out[b:b + 1], size=(100, 140), mode="bilinear", align_corners=False)
            o = o.sigmoid().numpy().squeeze()
            o = (o - o.min()) / (o.max() - o.min() + 1e-8)
            want.append((o * 255).astype(np.uint8))
        _check_u8(got, np.stack(want))
    assert len(inf._graphs) == 1 and inf.pv2_launches_per_step > 50
