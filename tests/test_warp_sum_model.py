"""A numpy model of `warp_sum_n` (csrc/vnl_device.cuh): the reduce-scatter + all-gather that replaces N butterfly reductions in the
solver's line search.  The claim the kernel relies on -- every lane ends with all N totals, BIT-IDENTICAL to N calls of the 5-stage
xor butterfly `warp_sum` (so the change could not move any parity result) -- is checked here in float32 arithmetic, lane by lane,
and the shuffle count is the one quoted in the source (2 N + 1 ... : 17 for N = 8, 10 for N = 4, instead of 5 N)."""
import numpy as np
import pytest


def butterfly(v):  # v [32] float32 -> every lane's result of `for o in 16, 8, 4, 2, 1: v += shfl_xor(v, o)`
    v = v.astype(np.float32).copy()
    lanes = np.arange(32)
    for o in (16, 8, 4, 2, 1):
        v = (v + v[lanes ^ o]).astype(np.float32)
    return v


def warp_sum_n(vals):  # vals [N][32] float32 -> ([N][32] results, number of shuffles), the device code statement by statement
    N = len(vals)
    LOG = {2: 1, 4: 2, 8: 3}[N]
    v = [x.astype(np.float32).copy() for x in vals]
    lanes = np.arange(32)
    shuffles, o, n = 0, 16, N
    while n > 1:
        up = (lanes & o) != 0
        for k in range(n // 2):
            send = np.where(up, v[k], v[k + n // 2])
            keep = np.where(up, v[k + n // 2], v[k])
            v[k] = (keep + send[lanes ^ o]).astype(np.float32)
            shuffles += 1
        n //= 2
        o //= 2
    t = v[0]
    q = 16 >> LOG
    while q > 0:
        t = (t + t[lanes ^ q]).astype(np.float32)
        shuffles += 1
        q //= 2
    out = []
    for j in range(N):
        out.append(np.full(32, t[j << (5 - LOG)], dtype=np.float32))
        shuffles += 1
    return out, shuffles


@pytest.mark.parametrize("N,count", [(2, 7), (4, 10), (8, 17)])
def test_warp_sum_n_is_bit_identical_to_the_butterflies(N, count):
    rng = np.random.default_rng(N)
    for trial in range(50):
        scale = 10.0 ** rng.integers(-6, 6)
        vals = [(scale * rng.standard_normal(32) * 10.0 ** rng.integers(-3, 3, 32)).astype(np.float32) for _ in range(N)]
        got, shuffles = warp_sum_n(vals)
        assert shuffles == count
        for j in range(N):
            want = butterfly(vals[j])
            assert np.all(want == want[0])  # the butterfly leaves the same bits on every lane
            assert np.array_equal(got[j].view(np.uint32), want.view(np.uint32)), (N, trial, j)
