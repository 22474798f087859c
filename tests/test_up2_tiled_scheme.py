"""Window arithmetic of up2_nhwc_bwd4_tiled_kernel (pranet-v2_b200/csrc/layout_bn.cu) restated with float32 numpy: for every plane size
the x2 align-corners backward can meet, the output rows / columns an 8-pixel input tile touches (a) fit the 22-wide shared-memory window the
kernel stages, and (b) contain every output pixel whose taps reach the tile, so no contribution is dropped."""
import numpy as np
import pytest

f32 = np.float32
UT, UWIN = 8, 22


def ratio(n):
    return f32(n - 1) / f32(2 * n - 1)


def lo_of(i, r):
    return max(0, int(np.floor(f32(i - 1) / r))) if r > 0 else 0


def hi_of(i, r, on):
    return min(on - 1, int(np.ceil(f32(i + 1) / r))) if r > 0 else on - 1


def taps(o, n, r):
    """pv2::bilinear_tap, align_corners=True: source rows of output row o."""
    src = r * f32(o)
    i0 = min(int(src), n - 1)
    i1 = i0 + (1 if i0 < n - 1 else 0)
    w1 = min(max(f32(src) - f32(i0), f32(0)), f32(1))
    return (i0, f32(1) - w1), (i1, w1)


@pytest.mark.parametrize("n", list(range(2, 70)) + [88, 96, 128, 176, 352])
def test_tile_windows(n):
    r, on = ratio(n), 2 * n
    touching = {i: [] for i in range(n)}
    for o in range(on):
        for i, w in taps(o, n, r):
            if w != 0:
                touching[i].append(o)
    for i0 in range(0, n, UT):
        i1 = min(i0 + UT, n) - 1
        w0, w1 = lo_of(i0, r), hi_of(i1, r, on)
        assert w1 - w0 + 1 <= UWIN, (n, i0, w0, w1)
        for i in range(i0, i1 + 1):
            lo, hi = lo_of(i, r), hi_of(i, r, on)
            assert w0 <= lo and hi <= w1
            assert hi - lo + 1 <= 7, (n, i, lo, hi)                     # the kernel's NC = 7 candidates per axis
            assert all(lo <= o <= hi for o in touching[i]), (n, i)     # every contributing output row is a candidate
