"""CPU model of the tiling / fold scheme of structure_loss_lowres_bwd_kernel + lowres_grad_fold_kernel
(pranet-v2_b200/csrc/structure_loss.cu): the same tile spans, quad pre-reduction (3 source columns per quad), lane ranges of the
x fold, slot layout and candidate-tile ranges of the fold kernel, restated with numpy float32 index math, must reproduce the
transpose of F.interpolate(mode='bilinear', align_corners=False) for every geometry the C ABI accepts (ratios <= 1/4, W % 4 == 0).
This pins the index logic on CPU; the kernels themselves are checked on the GPU in test_gpu_ops.py."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

f32 = np.float32
FT_H, FT_W, RMAX, CMAX = 32, 128, 10, 34


def tap(o, in_size, ratio):
    """pv2::bilinear_tap, align_corners=False."""
    src = max(f32(ratio) * (f32(o) + f32(0.5)) - f32(0.5), f32(0.0))
    i0 = min(int(src), in_size - 1)
    i1 = i0 + (1 if i0 < in_size - 1 else 0)
    w1 = min(max(f32(src) - f32(i0), f32(0.0)), f32(1.0))
    return i0, i1, f32(1.0) - w1, w1


def tile_span(o0, n, out_size, in_size, ratio):
    first = tap(min(o0, out_size - 1), in_size, ratio)[0]
    count = tap(min(o0 + n - 1, out_size - 1), in_size, ratio)[1] - first + 1
    return first, count


def model_backward(g, ih, iw, rh, rw):
    """g: (H, W) float32 gradient of the upsampled map -> (ih, iw) gradient of the low-res map, through the kernel's scheme."""
    H, W = g.shape
    tiles_x, tiles_y = (W + FT_W - 1) // FT_W, (H + FT_H - 1) // FT_H
    slots = {}
    for tile in range(tiles_x * tiles_y):
        y0, x0 = (tile // tiles_x) * FT_H, (tile % tiles_x) * FT_W
        xt = [tap(min(x0 + c, W - 1), iw, rw) for c in range(FT_W)]
        yt = [tap(min(y0 + r, H - 1), ih, rh) for r in range(FT_H)]
        r_first, nrows = tile_span(y0, FT_H, H, ih, rh)
        c_first, ncols = tile_span(x0, FT_W, W, iw, rw)
        assert nrows <= RMAX and ncols <= CMAX, (nrows, ncols)
        # A. quad pre-reduction
        Q = np.zeros((FT_H, 32, 3), np.float64)
        for row in range(FT_H):
            gy = y0 + row
            for lane in range(32):
                gx = x0 + 4 * lane
                if gy >= H or gx >= W:
                    continue
                cb = xt[4 * lane][0]
                for j in range(4):
                    i0, _, w0, w1 = xt[4 * lane + j]
                    d0 = i0 - cb
                    assert d0 in (0, 1)
                    e1 = d0 + (1 if cb + d0 < iw - 1 else 0)
                    assert e1 <= 2
                    Q[row, lane, d0] += g[gy, gx + j] * w0
                    Q[row, lane, e1] += g[gy, gx + j] * w1
        # B. x fold over the conservative lane range
        Hs = np.zeros((ncols, FT_H), np.float64)
        inv_r = f32(1.0) / f32(rw)
        for cl in range(ncols):
            c = c_first + cl
            l_lo = int(math.floor(float(((f32(c - 2) + f32(0.5)) * inv_r - f32(0.5) - f32(x0)) * f32(0.25)))) - 1
            l_hi = int(math.ceil(float(((f32(c + 1) + f32(0.5)) * inv_r - f32(0.5) - f32(x0)) * f32(0.25)))) + 1
            l_lo, l_hi = max(l_lo, 0), min(l_hi, 31)
            if c == 0:
                l_lo = 0
            for lane in range(32):
                d = c - xt[4 * lane][0]
                if 0 <= d <= 2 and np.any(Q[:, lane, d] != 0):
                    assert l_lo <= lane <= l_hi, ("lane range misses a contributor", c, lane, l_lo, l_hi)
            for lane in range(l_lo, l_hi + 1):
                d = c - xt[4 * lane][0]
                if 0 <= d <= 2:
                    Hs[cl] += Q[:, lane, d]
        # C. y fold into the tile's slot
        slot = np.zeros((RMAX, CMAX), np.float64)
        for rl in range(nrows):
            r = r_first + rl
            for row in range(FT_H):
                i0, _, w0, w1 = yt[row]
                i1 = min(i0 + 1, ih - 1)
                wy = (w0 if i0 == r else 0.0) + (w1 if i1 == r else 0.0)
                if wy:
                    slot[rl, :ncols] += wy * Hs[:, row]
        slots[tile] = slot
    # fold kernel
    out = np.zeros((ih, iw), np.float64)
    irh, irw = f32(1.0) / f32(rh), f32(1.0) / f32(rw)
    for r in range(ih):
        oy_lo = int(math.floor(float((f32(r) - f32(0.5)) * irh - f32(0.5)))) - 1
        oy_hi = int(math.ceil(float((f32(r) + f32(1.5)) * irh - f32(0.5)))) + 1
        if r == 0:
            oy_lo = 0
        if r == ih - 1:
            oy_hi = H - 1
        ty_lo, ty_hi = max(oy_lo, 0) // FT_H, min(min(oy_hi, H - 1) // FT_H, tiles_y - 1)
        for c in range(iw):
            ox_lo = int(math.floor(float((f32(c) - f32(0.5)) * irw - f32(0.5)))) - 1
            ox_hi = int(math.ceil(float((f32(c) + f32(1.5)) * irw - f32(0.5)))) + 1
            if c == 0:
                ox_lo = 0
            if c == iw - 1:
                ox_hi = W - 1
            tx_lo, tx_hi = max(ox_lo, 0) // FT_W, min(min(ox_hi, W - 1) // FT_W, tiles_x - 1)
            acc, seen = 0.0, set()
            for tyi in range(ty_lo, ty_hi + 1):
                r_first, nrows = tile_span(tyi * FT_H, FT_H, H, ih, rh)
                rl = r - r_first
                if rl < 0 or rl >= nrows:
                    continue
                for txi in range(tx_lo, tx_hi + 1):
                    c_first, ncols = tile_span(txi * FT_W, FT_W, W, iw, rw)
                    cl = c - c_first
                    if cl < 0 or cl >= ncols:
                        continue
                    acc += slots[tyi * tiles_x + txi][rl, cl]
                    seen.add(tyi * tiles_x + txi)
            # every tile holding a non-zero entry for (r, c) must have been visited
            for tile, slot in slots.items():
                if tile in seen:
                    continue
                rf, nr = tile_span((tile // tiles_x) * FT_H, FT_H, H, ih, rh)
                cf, nc = tile_span((tile % tiles_x) * FT_W, FT_W, W, iw, rw)
                if 0 <= r - rf < nr and 0 <= c - cf < nc:
                    assert slot[r - rf, c - cf] == 0.0, ("fold kernel misses a tile", r, c, tile)
            out[r, c] = acc
    return out


# (H, W, ih, iw, scale_factor or None for the size= form)
CASES = [
    (352, 352, 44, 44, 8), (352, 352, 22, 22, 16), (352, 352, 11, 11, 32),      # PraNet-V2 at 352^2 (pranet.py:349-415)
    (224, 224, 56, 56, 4), (224, 224, 7, 7, 32),                                # EMCAD (networks.py:116-123)
    (256, 256, 32, 32, 8), (448, 448, 14, 14, 32),                              # multi-scale training sizes
    (96, 132, 12, 33, None), (64, 100, 7, 9, None), (40, 260, 10, 65, 4),       # ragged: non-square, W % 128 != 0, non-integer ratios
]


@pytest.mark.parametrize("H,W,ih,iw,sf", CASES)
def test_lowres_backward_scheme_matches_interpolate_transpose(H, W, ih, iw, sf):
    rng = np.random.default_rng(H * 1000 + W + ih)
    g = rng.standard_normal((H, W)).astype(np.float32)
    if sf is not None:
        assert (int(math.floor(ih * sf)), int(math.floor(iw * sf))) == (H, W)
        rh = rw = float(f32(1.0 / sf))
        kw = dict(scale_factor=sf)
    else:
        rh, rw = float(f32(ih) / f32(H)), float(f32(iw) / f32(W))
        kw = dict(size=(H, W))
    assert rh <= 0.25 and rw <= 0.25 and W % 4 == 0
    low = torch.zeros(1, 1, ih, iw, dtype=torch.float64, requires_grad=True)
    F.interpolate(low, mode="bilinear", align_corners=False, **kw).backward(torch.from_numpy(g).double()[None, None])
    ref = low.grad[0, 0].numpy()
    got = model_backward(g, ih, iw, rh, rw)
    assert np.abs(got - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())
