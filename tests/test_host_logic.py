"""CPU tests of the host-side finishing logic (partitionsolvers_b200/host.py, the Python mirror of the
C++ facade): given the engine's raw outputs (sort index, boundaries, per-subset scores) it must
reproduce the reference's subsets, reorder, regulariser, getter quirk, sweep table and OLS choice.
The raw outputs are synthesised from the golden fixtures, so no GPU is needed."""
import numpy as np
import pytest

from conftest import assert_same_partition, golden_names, load_golden
from partitionsolvers_b200 import host


def _raw_from_golden(g, oracle):
    """Walk the fixture's nextStart_ rows like the device backtrace does."""
    dt = g["dtype"]
    ns, perm = g["next_start"], g["perm"]
    n = len(perm)
    Tc = ns.shape[0] - 1
    a_s, b_s = g["a"][perm], g["b"][perm]
    sweep = bool(g["sweep"])
    nS = Tc + 1 if sweep else 1
    bounds = -np.ones((nS, Tc + 1), np.int32)
    sc = np.zeros((nS, Tc), dt)
    for row in range(nS):
        S = row if sweep else Tc
        cur = 0
        bounds[row, 0] = 0
        for t in range(S, 0, -1):
            nxt = int(ns[t, cur])
            bounds[row, S - t + 1] = nxt
            sc[row, S - t] = oracle.range_score(a_s, b_s, int(g["dist"]), bool(g["rp"]), cur, nxt, dtype=dt)
            cur = nxt
    return dict(n=n, T=Tc, R=1, perm=perm, bounds=bounds, subset_scores=sc, max_score=None, next_start=None)


class _FakeCtx:
    def __init__(self, raw):
        self.raw = raw

    def solve_raw(self, *a, **k):
        return self.raw


@pytest.mark.parametrize("name", golden_names())
def test_facade_finishing_matches_reference(oracle, name):
    g = load_golden(name)
    dt = g["dtype"]
    raw = _raw_from_golden(g, oracle)
    s = host.DPSolver(len(g["a"]), int(g["T"]), g["a"], g["b"], int(g["dist"]), bool(g["rp"]), True,
                      float(g["gamma"]), int(g["reg_power"]), bool(g["sweep"]), False, dtype=dt, ctx=_FakeCtx(raw))
    assert_same_partition(s.subsets, g["subsets"])
    assert np.array_equal(s.get_score_by_subset_extern(), g["score_by_subset"])
    assert s.get_optimal_score_extern() == g["score"]
    assert dt(s.optimal_score) == g["optimal_score_member"]
    if int(g["sweep"]):
        allp = s.all
        assert len(allp) == len(g["all"])
        for (s1, v1), (s2, v2) in zip(allp, g["all"]):
            assert v1 == v2
            assert_same_partition(s1, s2)


def test_ols_matches_oracle_restatement(oracle):
    rs = np.random.RandomState(5)
    for dt in (np.float32, np.float64):
        for _ in range(50):
            T = int(rs.randint(2, 30))
            sc = np.concatenate([[0.0], np.cumsum(rs.uniform(0.01, 1.0, T) * np.arange(T, 0, -1) ** rs.uniform(0, 3))])
            assert host.ols_num_clusters(sc.astype(dt), dt) == oracle.ols(sc, dtype=dt)


def test_ols_degenerate_scores_default_to_one(oracle):
    # non-increasing scores -> log of non-positive differences -> NaN/-inf everywhere -> index 0 -> t = 1
    for dt in (np.float32, np.float64):
        sc = np.array([0, 5, 5, 5, 5], dt)
        assert oracle.ols(sc, dtype=dt) == host.ols_num_clusters(sc, dt)


def test_ols_single_sample_is_one(oracle):
    """T = 2: one regression sample, an under-determined system.  Whatever solution is taken (Eigen's colPivHouseholderQr
    returns the basic one) the only candidate index is 0, so the choice is 1 -- for the facade's restatement and the oracle's."""
    for dt in (np.float32, np.float64):
        for sc in ([0.0, 3.0, 5.0], [0.0, 1.0, 1.5], [0.0, 2.0, 1.0]):
            sc = np.array(sc, dt)
            assert host.ols_num_clusters(sc, dt) == 1 and oracle.ols(sc, dtype=dt) == 1


def test_poisson_risk_part_score_is_quasi_convex_along_a_row():
    """The claim behind the range tiers of Poisson risk-partitioning (csrc/screen.cuh, ScreenUB<DIST_POISSON, true>::hull):
    with the elements in ascending priority a/b, f(k) = A ln(A/B) over the range [i, k) falls, then rises, as k grows -- so
    any k-range is bounded by the larger of its two end cells.  Checked here in extended precision on random data of the
    kinds the engine sees (counts, null-model counts, real-valued), for many rows."""
    rs = np.random.RandomState(77)
    for kind in range(3):
        n = 3000
        if kind == 0:
            b = rs.poisson(100, n).clip(1).astype(np.float64)
            a = rs.poisson(np.repeat(rs.uniform(0.2, 2.5, 10), n // 10) * b).clip(1).astype(np.float64)
        elif kind == 1:
            b = rs.poisson(100, n).clip(1).astype(np.float64)
            a = rs.poisson(b).clip(1).astype(np.float64)
        else:
            b = rs.uniform(1e-3, 10.0, n)
            a = rs.uniform(1e-3, 10.0, n)
        order = np.argsort(a / b, kind="stable")
        a, b = a[order].astype(np.longdouble), b[order].astype(np.longdouble)
        for i in rs.randint(0, n - 2, 60):
            A = np.cumsum(a[i:])
            B = np.cumsum(b[i:])
            f = A * np.log(A / B)
            d = np.diff(f)
            tol = 1e-12 * np.maximum(np.abs(f[1:]), np.abs(f[:-1])) + 1e-15
            rising = np.nonzero(d > tol)[0]
            if rising.size:  # after the first rise nothing falls any more
                assert not np.any(d[rising[0]:] < -tol[rising[0]:]), (kind, int(i))
            # hence every k-range is bounded by its end cells
            for lo, hi in ((0, len(f) - 1), (len(f) // 3, 2 * len(f) // 3)):
                if hi > lo:
                    assert f[lo:hi + 1].max() <= max(f[lo], f[hi]) + 1e-9 * max(abs(f[lo]), abs(f[hi]), 1.0)
