"""Index logic of pv2's multi-value warp reduction (csrc/mc_dual_loss.cu: multi_step / multi_warp_sum / multi_slot), restated in
Python and checked on CPU for every value count the kernels can instantiate (2 + 2C for C = 2..12, and all of 1..32): after the five
halving steps every value's warp total sits in exactly one lane, the lane multi_slot() names, and padding lanes are marked -1 (two
lanes adding to the same shared-memory slot would lose an update)."""
import random

import pytest


def multi_step(vals, n, off):
    """One halving step over the 32 lanes: a lane keeps the lower (bit clear) or upper (bit set) half of its n values, sends the
    other half to lane ^ off and adds what it receives; the upper half of an odd n is padded with a zero."""
    h = (n + 1) // 2
    out = []
    for lane in range(32):
        up, partner = bool(lane & off), vals[lane ^ off]
        pup = bool((lane ^ off) & off)
        row = []
        for j in range(h):
            lo, hi = vals[lane][j], (vals[lane][j + h] if j + h < n else 0.0)
            plo, phi = partner[j], (partner[j + h] if j + h < n else 0.0)
            row.append((hi if up else lo) + (plo if pup else phi))
        out.append(row)
    return out, h


def multi_slot(nv, lane):
    n, real, idx = nv, nv, 0
    for off in (16, 8, 4, 2, 1):
        h = (n + 1) // 2
        if lane & off:
            idx += h
            real = real - h if real > h else 0
        else:
            real = min(real, h)
        n = h
    return idx if real >= 1 else -1


@pytest.mark.parametrize("nv", list(range(1, 33)))
def test_every_total_lands_in_exactly_one_lane(nv):
    rng = random.Random(nv)
    vals = [[rng.uniform(-1, 1) for _ in range(nv)] for _ in range(32)]
    want = [sum(vals[lane][i] for lane in range(32)) for i in range(nv)]
    cur, n = vals, nv
    for off in (16, 8, 4, 2, 1):
        cur, n = multi_step(cur, n, off)
    assert n == 1
    owner = {}
    for lane in range(32):
        s = multi_slot(nv, lane)
        assert -1 <= s < nv
        if s >= 0:
            assert s not in owner, f"value {s} claimed by lanes {owner[s]} and {lane}"
            owner[s] = lane
    assert sorted(owner) == list(range(nv))
    for i, lane in owner.items():
        assert abs(cur[lane][0] - want[i]) <= 1e-9


def test_shuffle_count():
    """The point of it: 21 shuffles for the 20 sums of a C = 9 subset instead of 5 per value."""
    n, total = 20, 0
    for _ in range(5):
        n = (n + 1) // 2
        total += n
    assert total == 21
