"""CPU tests: the oracle port (oracle/dp_oracle.cpp) against the reference's golden vector, against
fixtures written by the unmodified reference (tools/make_golden.py), against the reference itself
where /root/reference is present, and against the properties the reference's own gtest suite checks
(src/cpp/test/gtest_all.cpp; names kept)."""
import numpy as np
import pytest

from conftest import assert_same_partition, golden_names, load_golden
from partitionsolvers_b200 import synth

# gtest_all.cpp:392-398
TESTBASELINES_EXPECTED = [
    [1, 2, 3, 4, 6, 8, 10, 12, 16, 19, 21, 22, 23, 25, 28, 29, 32, 35, 38],
    [13, 24],
    [0, 5, 7, 9, 11, 14, 15, 18, 27, 30, 33, 36, 37, 39],
    [17, 26],
    [20, 31, 34],
]


def test_TestBaselines_golden_vector(oracle):
    g = load_golden("testbaselines_gauss_rp_f32")
    r = oracle.solve("port", g["a"], g["b"], 5, 0, True, dtype=np.float32)
    assert [sorted(s.tolist()) for s in r["subsets"]] == TESTBASELINES_EXPECTED
    # the fixture (written by the reference) says the same
    assert [sorted(s.tolist()) for s in g["subsets"]] == TESTBASELINES_EXPECTED
    # second, unasserted case of the reference test (gtest_all.cpp:400-406) as the reference computes it
    g1 = load_golden("testbaselines_poisson_mc_f32")
    r1 = oracle.solve("port", g1["a"], g1["b"], 3, 1, False, dtype=np.float32)
    assert_same_partition(r1["subsets"], g1["subsets"])
    assert [s.tolist() for s in r1["subsets"]] == [[0, 1, 2, 3], [4, 5, 6], [7]]


@pytest.mark.parametrize("name", golden_names())
def test_port_matches_reference_fixture_bit_for_bit(oracle, name):
    g = load_golden(name)
    dt = g["dtype"]
    r = oracle.solve("port", g["a"], g["b"], int(g["T"]), int(g["dist"]), bool(g["rp"]),
                     gamma=float(g["gamma"]), reg_power=int(g["reg_power"]), sweep=bool(g["sweep"]), dtype=dt)
    assert np.array_equal(r["perm"], g["perm"])
    assert np.array_equal(r["max_score"], g["max_score"], equal_nan=True)
    assert np.array_equal(r["next_start"], g["next_start"])
    assert_same_partition(r["subsets"], g["subsets"])
    assert np.array_equal(r["score_by_subset"], g["score_by_subset"])
    assert r["score"] == g["score"]
    assert r["optimal_score_member"] == g["optimal_score_member"]
    if int(g["sweep"]):
        assert len(r["all"]) == len(g["all"])
        for (s1, v1), (s2, v2) in zip(r["all"], g["all"]):
            assert v1 == v2
            assert_same_partition(s1, s2)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("dist", [0, 1, 2])
@pytest.mark.parametrize("rp", [False, True])
def test_port_matches_live_reference(oracle, dtype, dist, rp):
    if not oracle.have("ref"):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    rs = np.random.RandomState(1234 + dist * 2 + rp)
    for trial in range(3):
        n = int(rs.randint(3, 400))
        T = int(rs.randint(1, 12))
        a = rs.uniform(0.5, 10, n).astype(dtype) if dist == 1 else rs.uniform(-10, 10, n).astype(dtype)
        b = rs.uniform(1e-3, 10, n).astype(dtype)
        if trial == 2:  # massive priority ties: a in {1,2}, b = 1 (gtest_all.cpp:924-958)
            a = rs.randint(1, 3, n).astype(dtype)
            b = np.ones(n, dtype)
        r1 = oracle.solve("ref", a, b, T, dist, rp, gamma=0.1, reg_power=2, dtype=dtype, sweep=bool(trial == 1))
        r2 = oracle.solve("port", a, b, T, dist, rp, gamma=0.1, reg_power=2, dtype=dtype, sweep=bool(trial == 1))
        assert np.array_equal(r1["perm"], r2["perm"])
        assert np.array_equal(r1["max_score"], r2["max_score"], equal_nan=True)
        assert np.array_equal(r1["next_start"], r2["next_start"])
        assert_same_partition(r1["subsets"], r2["subsets"])
        assert r1["score"] == r2["score"] or (np.isnan(r1["score"]) and np.isnan(r2["score"]))


def test_range_score_matches_reference_table(oracle):
    if not oracle.have("ref"):
        pytest.skip("oracle/_ref not built")
    rs = np.random.RandomState(7)
    n = 200
    for dtype in (np.float32, np.float64):
        a = rs.uniform(0.5, 10, n).astype(dtype)
        b = rs.uniform(0.5, 10, n).astype(dtype)
        ii = rs.randint(0, n - 1, 64)
        jj = np.array([rs.randint(i + 1, n + 1) for i in ii])
        for dist in (0, 1, 2):
            for rp in (False, True):
                want = oracle.ref_score_pairs(a, b, dist, rp, ii, jj, dtype=dtype)
                got = np.array([oracle.range_score(a, b, dist, rp, int(i), int(j), dtype=dtype)
                                for i, j in zip(ii, jj)])
                assert np.array_equal(got, want)


# ---- properties the reference's gtest suite checks, run on the port -------------------------------
def _sorted_inputs(n, seed, dist, dtype=np.float32):
    rs = np.random.RandomState(seed)
    a = rs.uniform(0.1, 10, n) if dist == 1 else rs.uniform(-10, 10, n)
    b = rs.uniform(0.1, 10, n)
    o = np.argsort(a / b, kind="stable")
    return a[o].astype(dtype), b[o].astype(dtype)


@pytest.mark.parametrize("dist", [0, 1, 2])
def test_TestOrderedProperty(oracle, dist):  # gtest_all.cpp:421-459
    a, b = _sorted_inputs(100, 3, dist)
    r = oracle.solve("port", a, b, 8, dist, False, dtype=np.float32)
    for s in r["subsets"]:
        v = np.sort(s)
        assert np.array_equal(v, np.arange(v[0], v[0] + len(v)))


@pytest.mark.parametrize("dist", [0, 1, 2])
def test_TestSinglePartitionSpecification(oracle, dist):  # gtest_all.cpp:461-498
    a, b = _sorted_inputs(50, 4, dist)
    r = oracle.solve("port", a, b, 1, dist, True, dtype=np.float32)
    assert len(r["subsets"]) == 1 and sorted(r["subsets"][0].tolist()) == list(range(50))


@pytest.mark.parametrize("dist", [0, 1])
def test_TestHighestScoringSetOf2TieOutAllDists(oracle, dist):  # gtest_all.cpp:500-559
    rs = np.random.RandomState(5 + dist)
    n = 300
    a = rs.uniform(0.1, 10, n).astype(np.float32)
    b = rs.uniform(0.1, 10, n).astype(np.float32)
    dp = oracle.solve("port", a, b, 2, dist, False, dtype=np.float32)
    lt = oracle.ltss("port", a, b, dist)
    assert sorted(dp["subsets"][1].tolist()) == sorted(lt["subset"].tolist())


@pytest.mark.parametrize("dist", [0, 1, 2])
def test_TestSweepResultsMatchSingleTResults(oracle, dist):  # gtest_all.cpp:603-662
    rs = np.random.RandomState(11 + dist)
    n, T = 120, 9
    a = rs.uniform(0.1, 10, n).astype(np.float32)
    b = rs.uniform(0.1, 10, n).astype(np.float32)
    sw = oracle.solve("port", a, b, T, dist, True, dtype=np.float32, sweep=True)
    for S in range(1, T + 1):
        one = oracle.solve("port", a, b, S, dist, True, dtype=np.float32)
        assert one["optimal_score_member"] == sw["all"][S][1]
        assert_same_partition(one["subsets"], sw["all"][S][0])


def test_TestBestSweepResultIsAlwaysMostGranularPartition(oracle):  # gtest_all.cpp:561-601
    rs = np.random.RandomState(13)
    n = 40
    a = rs.uniform(0.1, 10, n).astype(np.float64)
    b = rs.uniform(0.1, 10, n).astype(np.float64)
    sw = oracle.solve("port", a, b, n, 2, True, dtype=np.float64, sweep=True)
    sc = [float(v) for _, v in sw["all"]][1:]
    assert all(y >= x - 1e-9 * abs(x) for x, y in zip(sc, sc[1:]))


@pytest.mark.parametrize("dist", [0, 1, 2])
def test_TestOptimalityWithRandomPartitions(oracle, dist):  # gtest_all.cpp:960-1029
    rs = np.random.RandomState(17 + dist)
    n = 60
    a = rs.uniform(0.1, 10, n).astype(np.float64)
    b = rs.uniform(0.1, 10, n).astype(np.float64)
    r = oracle.solve("port", a, b, 3, dist, True, dtype=np.float64)
    a_s, b_s = a[r["perm"]], b[r["perm"]]
    best = float(r["optimal_score_member"])
    for _ in range(300):
        c = np.sort(rs.choice(np.arange(1, n), 2, replace=False))
        tot = sum(float(oracle.range_score(a_s, b_s, dist, True, lo, hi, dtype=np.float64))
                  for lo, hi in ((0, c[0]), (c[0], c[1]), (c[1], n)))
        assert tot <= best * (1 + 1e-12) + 1e-12


def test_TestConsistencyofOLSReturnedPartitionSize(oracle):  # gtest_all.cpp:664-733 (structure only)
    a, b = synth.poisson_planted(800, 4, seed=99, dtype=np.float32)
    r = oracle.solve("port", a, b, 10, 1, False, dtype=np.float32, sweep=True, ols=True)
    assert 1 <= r["ols_t"] <= 10
    assert len(r["subsets"]) == r["ols_t"]
    assert len(r["all"]) == 11 and all(len(r["all"][S][0]) == S for S in range(11))


def test_ols_recovers_planted_levels(oracle):  # solver_tests.py:181-201 shape: 4 planted levels
    rs = np.random.RandomState(0xC0FFEE3 % (2 ** 31))
    n = 2000
    b = rs.uniform(90, 110, n)
    q = np.repeat([0.25, 0.75, 1.25, 1.75], n // 4)
    a = rs.normal(q * b, 1.0)
    r = oracle.solve("port", a, b, 12, 2, True, dtype=np.float32, sweep=True, ols=True)
    assert r["ols_t"] == 4


def test_work_counts():
    assert synth.actual_cells(10, 1) == 0
    assert synth.actual_cells(10, 2) == 9
    # brute force
    n, T = 37, 6
    cnt = 0
    for j in range(2, T + 1):
        for i in range(0, n - j + 1):
            cnt += (n - (j - 1)) - i
            if j == T:
                break
    assert synth.actual_cells(n, T) == cnt
