"""Seeded synthetic inputs and weights shared by the oracle, the tests and bench.py.

TEST INFRASTRUCTURE (oracle side).  Everything here is deterministic for a given
seed and torch build (CPU generators), so golden vectors only need to store
OUTPUTS: inputs and weights are regenerated.

Workload shapes follow SURVEY.md §8(d): backbone features of PraNet-V2
(`binary_seg/lib/pranet.py:337-341`: x2 B×512×S/8, x3 B×1024×S/16, x4 B×2048×S/32;
PVT variant 128/320/512, `pranet.py:158-160`), polyp-like ellipse masks, and the soft
masks produced by the multi-scale resize of `binary_seg/MyTrain_med.py:70-73`.
"""
from __future__ import annotations

import zlib

import numpy as np
import torch
import torch.nn.functional as F

RES2NET_CH = (512, 1024, 2048)
PVT_CH = (128, 320, 512)


def _gen(seed: int, key: str = "") -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((int(seed) * 1000003 + zlib.crc32(key.encode())) % (2**63 - 1))
    return g


def synth_state_dict(template: dict, seed: int = 0) -> dict:
    """Fill every entry of a state_dict-shaped template with seeded, non-degenerate values.

    Independent of module construction order (keyed by parameter NAME), so the reference
    module, the oracle and the CUDA module can all be given identical weights without
    committing them.  BN affine / running stats are perturbed away from (1, 0, 0, 1) so
    that eval-mode parity actually exercises them.
    """
    out = {}
    for k, t in template.items():
        g = _gen(seed, k)
        shape = tuple(t.shape)
        if k.endswith("num_batches_tracked"):
            v = torch.zeros(shape, dtype=torch.long)
        elif k.endswith("running_mean"):
            v = 0.1 * torch.randn(shape, generator=g)
        elif k.endswith("running_var"):
            v = 0.6 + 0.8 * torch.rand(shape, generator=g)
        elif len(shape) >= 2:  # conv / linear weight: variance-preserving
            fan_in = int(np.prod(shape[1:]))
            v = torch.randn(shape, generator=g) * (1.0 / np.sqrt(fan_in))
        elif k.endswith("weight"):  # norm gamma
            v = 1.0 + 0.2 * torch.randn(shape, generator=g)
        else:  # bias / beta
            v = 0.1 * torch.randn(shape, generator=g)
        out[k] = v.to(t.dtype) if t.dtype.is_floating_point else v
    return out


def backbone_features(batch: int, size: int, seed: int = 0, channels=RES2NET_CH):
    """relu(randn) pyramids of the shapes the head consumes."""
    feats = []
    for i, (c, s) in enumerate(zip(channels, (8, 16, 32))):
        g = _gen(seed, f"feat{i}")
        feats.append(torch.relu(torch.randn(batch, c, size // s, size // s, generator=g)))
    return feats


def ellipse_masks(batch: int, h: int, w: int, seed: int = 0) -> torch.Tensor:
    """Binary polyp-like masks: union of 1-3 filled ellipses (5-30 % of the frame)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    out = np.zeros((batch, 1, h, w), np.float32)
    for b in range(batch):
        for _ in range(int(rng.integers(1, 4))):
            cy, cx = rng.uniform(0.2, 0.8) * h, rng.uniform(0.2, 0.8) * w
            ry, rx = rng.uniform(0.08, 0.25) * h, rng.uniform(0.08, 0.25) * w
            th = rng.uniform(0, np.pi)
            dy, dx = yy - cy, xx - cx
            u = (dx * np.cos(th) + dy * np.sin(th)) / rx
            v = (-dx * np.sin(th) + dy * np.cos(th)) / ry
            out[b, 0][(u * u + v * v) <= 1.0] = 1.0
    return torch.from_numpy(out)


def soft_masks(batch: int, h: int, w: int, seed: int = 0, base: int | None = None) -> torch.Tensor:
    """Soft masks as the multi-scale training path makes them (MyTrain_med.py:72-73):
    binary masks at `base` resolution bilinearly resized with align_corners=True."""
    base = base or max(8, int(round(h / 1.25 / 8)) * 8)
    m = ellipse_masks(batch, base, base, seed)
    return F.interpolate(m, size=(h, w), mode="bilinear", align_corners=True)


def logits(shape, seed: int = 0, key: str = "logit", scale: float = 3.0) -> torch.Tensor:
    return scale * torch.randn(*shape, generator=_gen(seed, key))


def class_labels(batch: int, h: int, w: int, num_classes: int, seed: int = 0) -> torch.Tensor:
    """Synapse-shaped label maps: background 0 plus random ellipses for organs 1..C-1."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    out = np.zeros((batch, h, w), np.int64)
    for b in range(batch):
        for c in range(1, num_classes):
            cy, cx = rng.uniform(0.15, 0.85) * h, rng.uniform(0.15, 0.85) * w
            ry, rx = rng.uniform(0.04, 0.14) * h, rng.uniform(0.04, 0.14) * w
            out[b][((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0] = c
    return torch.from_numpy(out)
