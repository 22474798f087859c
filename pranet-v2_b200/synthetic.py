"""Synthetic workload generators for benchmarks (there is no dataset on the box): polyp-like binary masks
(union of 1-3 filled ellipses covering ~5-30 % of the frame) and random images of the training shape."""
from __future__ import annotations

import numpy as np
import torch


def ellipse_masks(batch: int, h: int, w: int, seed: int = 0) -> torch.Tensor:
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


def images(batch: int, size: int, seed: int = 0) -> torch.Tensor:
    return torch.randn(batch, 3, size, size, generator=torch.Generator().manual_seed(seed))
