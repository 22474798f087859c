"""CPU model of bilinear_bwd2_kernel (pranet-v2_b200/csrc/bilinear.cu): the separable backward of the exact x8 / x16 / x32 / x64
half-pixel up-scalings.  The kernel's scheduling -- bands of input rows per CTA, cells (= output rows that share their upper source
row), 4-row chunks, column quads that share their source column, the (a, b) x (lo, hi) partial sums and the order in which pass 2
adds them -- is restated with numpy float32 tap arithmetic and must reproduce the transpose of
F.interpolate(mode='bilinear', align_corners=False).  This pins the index logic (including the clamped first / last rows and
columns, ragged last bands, and the row caps) on CPU; the kernel itself is checked on the GPU in test_gpu_ops.py."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

f32 = np.float32
MAX_CHUNKS, SMEM = 52, 64 * 1024


def tap(o, in_size, ratio):
    """pv2::bilinear_tap, align_corners=False."""
    src = max(f32(ratio) * (f32(o) + f32(0.5)) - f32(0.5), f32(0.0))
    i0 = min(int(src), in_size - 1)
    w1 = min(max(f32(src) - f32(i0), f32(0.0)), f32(1.0))
    return i0, f32(1.0) - w1, w1


def rows_per_cta(ih, s, ow, bands):
    r = (ih + bands - 1) // bands
    cap = (MAX_CHUNKS - s // 8) * 4 // s - 1
    cap2 = SMEM // (4 * ow) - 1
    return max(1, min(r, cap, cap2))


def fast_path_ok(ih, iw, oh, ow):
    if ow % 4 or ow // 4 > 384 or oh % ih or ow % iw:
        return False
    s = oh // ih
    if s != ow // iw or s % 8 or s > 64 or iw > 96:
        return False
    R = rows_per_cta(ih, s, ow, 1)
    return (R + 1) * s // 4 + s // 8 <= MAX_CHUNKS and 4 * (R + 1) * ow <= SMEM


def model_backward(g, ih, iw, bands):
    """g: (oh, ow) float32 upstream gradient -> (ih, iw) gradient of the low-res plane, through the kernel's scheme."""
    oh, ow = g.shape
    s = oh // ih
    ratio = f32(1.0) / f32(s)
    R = rows_per_cta(ih, s, ow, bands)
    ow4 = ow // 4
    xt = [tap(x, iw, ratio) for x in range(ow)]
    quad_i0 = [xt[4 * q][0] for q in range(ow4)]
    for q in range(ow4):                        # the structural assumption: a quad has ONE source column
        assert all(xt[4 * q + e][0] == quad_i0[q] for e in range(4))
    run_q0 = [0] * (iw + 1)
    for k in range(ow4 + 1):
        prev = -1 if k == 0 else quad_i0[k - 1]
        cur = iw if k == ow4 else quad_i0[k]
        for c in range(prev + 1, cur + 1):
            run_q0[c] = k
    din = np.full((ih, iw), np.nan, np.float64)
    ya = 0
    while ya < ih:
        yb = min(ya + R, ih)
        c0 = max(ya - 1, 0)
        ncell = yb - c0
        row0 = 0 if c0 == 0 else c0 * s + s // 2
        row1 = oh if yb == ih else yb * s + s // 2
        assert (row1 - row0) % 4 == 0
        nchunk = (row1 - row0) // 4
        assert nchunk <= MAX_CHUNKS
        yt = [tap(row0 + j, ih, ratio) for j in range(nchunk * 4)]
        chunk_i0 = [yt[4 * k][0] for k in range(nchunk)]
        for k in range(nchunk):                 # a chunk has ONE upper source row
            assert all(yt[4 * k + u][0] == chunk_i0[k] for u in range(4))
        assert chunk_i0[0] == c0
        cell_k0 = [0] * (ncell + 1)
        for k in range(nchunk + 1):
            prev = c0 - 1 if k == 0 else chunk_i0[k - 1]
            cur = c0 + ncell if k == nchunk else chunk_i0[k]
            for c in range(prev + 1, cur + 1):
                cell_k0[c - c0] = k
        # pass 1: per (cell, quad): a = sum w0x * v, b = sum w1x * v folded with the rows' (w0y, w1y)
        part = np.zeros((4, ncell, ow4), np.float64)
        for cell in range(ncell):
            for k in range(cell_k0[cell], cell_k0[cell + 1]):
                for u in range(4):
                    _, wy0, wy1 = yt[4 * k + u]
                    row = g[row0 + 4 * k + u]
                    for q in range(ow4):
                        a = sum(float(xt[4 * q + e][1]) * float(row[4 * q + e]) for e in range(4))
                        b = sum(float(xt[4 * q + e][2]) * float(row[4 * q + e]) for e in range(4))
                        part[0, cell, q] += float(wy0) * a
                        part[1, cell, q] += float(wy1) * a
                        part[2, cell, q] += float(wy0) * b
                        part[3, cell, q] += float(wy1) * b
        # pass 2
        for y in range(ya, yb):
            for ix in range(iw):
                acc = 0.0

                def fold(cell, ysel):
                    t = 0.0
                    pa, pb = part[ysel, cell - c0], part[2 + ysel, cell - c0]
                    t += sum(pa[q] for q in range(run_q0[ix], run_q0[ix + 1]))
                    if ix > 0:
                        t += sum(pb[q] for q in range(run_q0[ix - 1], run_q0[ix]))
                    if ix == iw - 1:
                        t += sum(pb[q] for q in range(run_q0[ix], run_q0[ix + 1]))
                    return t
                acc += fold(y, 0)
                if y > 0:
                    acc += fold(y - 1, 1)
                if y == ih - 1:
                    acc += fold(y, 1)
                assert np.isnan(din[y, ix])
                din[y, ix] = acc
        ya = yb
    return din


@pytest.mark.parametrize("ih,iw,s,bands", [(11, 11, 8, 2), (11, 11, 32, 2), (6, 9, 16, 2), (13, 5, 8, 3), (3, 5, 64, 2), (1, 4, 8, 2), (12, 12, 8, 18),
                                            (5, 3, 32, 4), (2, 2, 16, 2)])
def test_model_matches_aten_transpose(ih, iw, s, bands):
    oh, ow = ih * s, iw * s
    assert fast_path_ok(ih, iw, oh, ow)
    rng = np.random.default_rng(ih * 100 + s)
    g = rng.standard_normal((oh, ow)).astype(np.float32)
    x = torch.zeros(1, 1, ih, iw, dtype=torch.float64, requires_grad=True)
    F.interpolate(x, scale_factor=s, mode="bilinear", align_corners=False).backward(torch.from_numpy(g).double()[None, None])
    want = x.grad[0, 0].numpy()
    got = model_backward(g, ih, iw, bands)
    assert np.isfinite(got).all()
    np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-5 * np.abs(want).max())


def test_fast_path_predicate():
    assert fast_path_ok(44, 44, 352, 352) and fast_path_ok(22, 22, 352, 352) and fast_path_ok(11, 11, 352, 352) and fast_path_ok(88, 88, 704, 704)
    assert not fast_path_ok(56, 56, 224, 224)        # x4: quads straddle two source columns
    assert not fast_path_ok(11, 11, 22, 22)          # x2
    assert not fast_path_ok(7, 7, 50, 61)            # fractional ratio
    # row caps: window rows (R + 1) * s + s / 2 fit the 52-chunk tables, partial sums fit 64 KB
    for ih, s, ow in ((44, 8, 352), (22, 16, 352), (11, 32, 352), (88, 8, 704), (3, 64, 192)):
        for bands in (1, 2, 5, 18):
            R = rows_per_cta(ih, s, ow, bands)
            assert R >= 1 and (R + 1) * s // 4 + s // 8 <= MAX_CHUNKS and 4 * (R + 1) * ow <= SMEM
