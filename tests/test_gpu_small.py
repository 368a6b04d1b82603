"""Small instances (ps_b200.cu: SMALL_ROWS geometry -- 32-row tiles of 32 cells, the warp-per-row-block level-1 walk, the
one-kernel bitonic sort, results staged through pinned memory): every table cell, the sort index, subsets and scores against
the oracle port, at the sizes where the geometry, the sort's power-of-two padding and the 128-row fallback change over."""
import numpy as np
import pytest

from partitionsolvers_b200 import capi, synth

pytestmark = pytest.mark.gpu

SIZES = [2, 31, 32, 33, 63, 64, 65, 127, 128, 129, 1000, 1023, 1024, 1025, 2089, 4095, 4096, 4097]


def _stable_order(a, b, dt):
    return np.argsort((np.asarray(a, dt) / np.asarray(b, dt)).astype(dt), kind="stable").astype(np.int32)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_small_sort_is_the_stable_order(ctx, dt):
    """Count data (many tied priorities) and real-valued data: the device's sort index is the stable argsort of a/b computed
    in DataType -- what the radix passes produce for larger n, so results do not depend on which sort ran."""
    rs = np.random.RandomState(5)
    for n in SIZES:
        for kind in ("counts", "real"):
            if kind == "counts":
                b = rs.poisson(20, n).clip(1).astype(dt)
                a = rs.poisson(b).clip(1).astype(dt)
            else:
                a, b = synth.uniform_ab(n, seed=100 + n, dtype=dt)
            raw = ctx.solve_raw(a, b, min(3, n), 2, False, dtype=dt)
            assert np.array_equal(raw["perm"], _stable_order(a, b, dt)), (n, kind)


@pytest.mark.parametrize("dist,rp", [(1, False), (1, True), (0, False), (0, True), (2, False)])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_small_geometry_tables_equal_the_oracle(oracle, ctx, dt, dist, rp):
    """RUNNING sums on real-valued data (every addition rounds: the order of the walk matters) and on counts."""
    for n in SIZES:
        if n > 1100 and (dist, rp) not in ((1, False), (2, False)):
            continue  # the oracle port is O(n^2 T) on one core: the larger sizes for two scores only
        T = min(5, n)
        if dist == 1:
            a, b = synth.poisson_planted(n, 4, seed=200 + n, dtype=dt) if n % 2 else synth.uniform_ab(n, seed=300 + n, dtype=dt)
            a = np.maximum(a, dt(0.25))
        else:
            a, b = synth.gaussian_mixture(n, 4, seed=400 + n, dtype=dt)
        ref = oracle.solve("port", a, b, T, dist, rp, dtype=dt)
        raw = ctx.solve_raw(a, b, T, dist, rp, dtype=dt, perm=ref["perm"], levels=True)
        assert np.array_equal(raw["max_score"], ref["max_score"]), (n, dist, rp)
        assert np.array_equal(raw["next_start"], ref["next_start"]), (n, dist, rp)
        # the 128-row geometry (a user-given chunk length switches the small geometry off) and PREFIX sums on the small one
        big = ctx.solve_raw(a, b, T, dist, rp, dtype=dt, perm=ref["perm"], levels=True, chunk_len=128)
        for k in ("max_score", "next_start", "bounds", "subset_scores"):
            assert np.array_equal(raw[k], big[k]), (n, k)


def test_small_batches_and_sweep(oracle, ctx):
    """A few small instances in one call (the geometry is chosen from n and the batch size) and the sweep table."""
    n, T, R = 700, 6, 5
    ab = [synth.poisson_planted(n, 3 + r, seed=500 + r, dtype=np.float32) for r in range(R)]
    a = np.stack([x[0] for x in ab])
    b = np.stack([x[1] for x in ab])
    raw = ctx.solve_raw(a, b, T, 1, True, dtype=np.float32, levels=True, sweep=True)
    for r in range(R):
        one = ctx.solve_raw(a[r], b[r], T, 1, True, dtype=np.float32, levels=True, sweep=True, chunk_len=128)
        for k in ("perm", "max_score", "next_start", "bounds", "subset_scores"):
            assert np.array_equal(raw[k][r], one[k]), (r, k)
