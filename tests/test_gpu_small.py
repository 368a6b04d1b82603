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


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_repeated_small_solves_replay_a_graph_with_the_same_results(oracle, dt):
    """The second solve of a shape is recorded as a CUDA graph and every later one replays it (ps_b200.cu:
    solve_small_graph): different data through the same graph must give what an eager context gives, tables included."""
    eager = capi.Context(0)
    g = capi.Context(0)
    n, T = 1500, 5
    launches = []
    for rep in range(5):
        for dist, rp in ((1, False), (2, False), (0, True)):
            a, b = synth.gaussian_mixture(n, 4, seed=700 + rep, dtype=dt)
            if dist == 1:
                a = np.maximum(a, dt(0.5))
            got = g.solve_raw(a, b, T, dist, rp, dtype=dt, levels=True, sweep=True)
            launches.append(g.timings()["kernel_launches"])
            # the eager side: a user-given chunk length is never recorded (and takes the 128-row geometry)
            want = eager.solve_raw(a, b, T, dist, rp, dtype=dt, levels=True, sweep=True, chunk_len=128)
            for k in ("perm", "max_score", "next_start", "bounds", "subset_scores"):
                assert np.array_equal(got[k], want[k]), (rep, dist, k)
            ref = oracle.solve("port", a, b, T, dist, rp, dtype=dt)
            if np.array_equal(got["perm"], ref["perm"]):  # tied priorities: the port's order may differ from the stable one
                assert np.array_equal(got["max_score"], ref["max_score"]), (rep, dist)
                assert np.array_equal(got["next_start"], ref["next_start"]), (rep, dist)
    assert min(launches) > 0


def test_graph_replay_falls_back_when_the_precheck_fails(oracle):
    """A recorded solve assumes the inputs pass the safe-range / positivity precheck; a replay on data that does not (a <= 0
    under the Poisson score, a value outside the safe range) must be redone on the general-math kernels."""
    g = capi.Context(0)
    n, T, dt = 900, 4, np.float64
    a, b = synth.uniform_ab(n, seed=11, dtype=dt)
    for _ in range(3):
        g.solve_raw(a, b, T, 2, False, dtype=dt, levels=True)  # records and replays
    a2 = a.copy()
    a2[10] = 1e-40  # outside the safe range of the branch-free math path
    got = g.solve_raw(a2, b, T, 2, False, dtype=dt, levels=True)
    ref = oracle.solve("port", a2, b, T, 2, False, dtype=dt)
    assert np.array_equal(got["max_score"], ref["max_score"]) and np.array_equal(got["next_start"], ref["next_start"])
    # ... and the graph is still good for the next well-behaved input
    got = g.solve_raw(a, b, T, 2, False, dtype=dt, levels=True)
    ref = oracle.solve("port", a, b, T, 2, False, dtype=dt)
    assert np.array_equal(got["max_score"], ref["max_score"]) and np.array_equal(got["next_start"], ref["next_start"])


def test_graphs_survive_workspace_growth():
    """A larger solve in between reallocates workspaces; recorded graphs hold the old pointers and must be retired."""
    g = capi.Context(0)
    a, b = synth.poisson_planted(800, 4, seed=21, dtype=np.float32)
    first = [g.solve_raw(a, b, 4, 1, True, dtype=np.float32, levels=True) for _ in range(3)]
    A, B = synth.poisson_planted(30000, 4, seed=22, dtype=np.float32)
    g.solve_raw(A, B, 4, 1, True, dtype=np.float32)
    again = [g.solve_raw(a, b, 4, 1, True, dtype=np.float32, levels=True) for _ in range(3)]
    for r in first + again:
        for k in ("perm", "max_score", "next_start", "bounds", "subset_scores"):
            assert np.array_equal(r[k], first[0][k]), k


def test_random_small_shapes_graph_replay_equals_eager():
    """Randomised: 40 shapes (n, T, score, objective, precision, sweep, batch), each solved three times with fresh data on a
    context that records / replays graphs and on one created with PS_NO_GRAPH=1 -- every output array must be identical."""
    import os
    os.environ["PS_NO_GRAPH"] = "1"
    try:
        eager = capi.Context(0)
    finally:
        del os.environ["PS_NO_GRAPH"]
    g = capi.Context(0)
    rs = np.random.RandomState(2024)
    for case in range(40):
        n = int(rs.choice([33, 64, 100, 257, 700, 1024, 1500, 2089, 3000, 4096]))
        T = int(rs.randint(1, 9))
        dist, rp = [(1, False), (1, True), (0, False), (0, True), (2, False)][rs.randint(5)]
        dt = np.float32 if rs.rand() < 0.5 else np.float64
        sweep = bool(rs.rand() < 0.4)
        R = int(rs.choice([1, 1, 1, 2, 3]))
        for rep in range(3):
            b = rs.uniform(0.5, 50.0, (R, n)) if rs.rand() < 0.5 else rs.poisson(30, (R, n)).clip(1).astype(np.float64)
            a = rs.uniform(0.5, 50.0, (R, n)) if dist == 1 else rs.normal(0.0, 5.0, (R, n)) + b
            a, b = a.astype(dt), b.astype(dt)
            if R == 1:
                a, b = a[0], b[0]
            x = g.solve_raw(a, b, T, dist, rp, dtype=dt, levels=True, sweep=sweep)
            y = eager.solve_raw(a, b, T, dist, rp, dtype=dt, levels=True, sweep=sweep)
            for k in ("perm", "max_score", "next_start", "bounds", "subset_scores"):
                assert np.array_equal(x[k], y[k]), (case, rep, n, T, dist, rp, dt.__name__, sweep, R, k)
