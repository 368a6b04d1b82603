"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI,
against the reference-written golden fixtures, against the oracle port on seeded inputs, and -- at
BASELINE.json's full sizes -- through size-independent properties plus oracle-checked sampled rows.

Tolerances (DESIGN.md "Parity"), RUNNING sums (the default):
  * every score, f32 and f64: bit-exact -- every maxScore_/nextStart_ cell, subsets, per-subset scores, objective.
    Gaussian/Rational need only IEEE arithmetic; Poisson additionally needs log/logf to carry the host libm's
    bits, which csrc/glibc_log.cuh provides (conftest.device_log_matches_host verifies it on the machine at hand;
    only if that probe fails do the Poisson checks fall back to 1e-9 (f64) / 2e-4 (f32) with tie-aware splits).
  * PREFIX sums: bit-identical to RUNNING when all partial sums are exact, else prefix-difference rounding.
"""
import numpy as np
import pytest

from conftest import assert_same_partition, check_levels, golden_names, load_golden, tol_for
from partitionsolvers_b200 import capi, host, synth

pytestmark = pytest.mark.gpu

TIE = {np.float32: 1e-5, np.float64: 1e-12}


def _recompute_fn(oracle, a_s, b_s, ms_o, dist, rp, dt):
    def f(j, i, k):
        return float(oracle.range_score(a_s, b_s, dist, rp, i, k, dtype=dt)) + float(ms_o[j - 1][k])
    return f


def _check_against(oracle, ctx, a, b, T, dist, rp, dt, ref, sum_mode=capi.PS_SUM_RUNNING, chunk_len=0,
                   exact_sums=True):
    """ref: dict with perm/max_score/next_start (fixture or oracle output)."""
    raw = ctx.solve_raw(a, b, T, dist, rp, dtype=dt, sum_mode=sum_mode, perm=ref["perm"], levels=True,
                        chunk_len=chunk_len)
    rtol = tol_for(dt, dist == 1)
    if sum_mode == capi.PS_SUM_PREFIX and not exact_sums:
        rtol = max(rtol, 1e-9 if dt == np.float64 else 5e-3)
    a_s, b_s = np.asarray(a, dt)[ref["perm"]], np.asarray(b, dt)[ref["perm"]]
    tie = TIE[dt] if rtol else 0.0
    if sum_mode == capi.PS_SUM_PREFIX and not exact_sums:
        tie = rtol
    flips = check_levels(raw["max_score"], raw["next_start"], ref["max_score"], ref["next_start"], rtol, tie,
                         _recompute_fn(oracle, a_s, b_s, ref["max_score"], dist, rp, dt))
    assert np.array_equal(raw["perm"], ref["perm"])
    return raw, flips


@pytest.mark.parametrize("name", golden_names())
def test_cuda_matches_reference_fixture(oracle, ctx, name):
    g = load_golden(name)
    dt = g["dtype"]
    T, dist, rp = int(g["T"]), int(g["dist"]), bool(g["rp"])
    raw, flips = _check_against(oracle, ctx, g["a"], g["b"], T, dist, rp, dt, g)
    s = host.DPSolver(len(g["a"]), T, g["a"], g["b"], dist, rp, True, float(g["gamma"]), int(g["reg_power"]),
                      bool(g["sweep"]), False, dtype=dt, ctx=ctx, perm=g["perm"])
    rtol = tol_for(dt, dist == 1)
    if rtol == 0.0:
        assert_same_partition(s.subsets, g["subsets"])
        assert np.array_equal(s.get_score_by_subset_extern(), g["score_by_subset"])
        assert s.get_optimal_score_extern() == g["score"]
        if int(g["sweep"]):
            for (s1, v1), (s2, v2) in zip(s.all, g["all"]):
                assert v1 == v2
                assert_same_partition(s1, s2)
    else:
        if flips == 0:
            assert_same_partition(s.subsets, g["subsets"])
        want = float(g["score"])
        assert abs(float(s.get_optimal_score_extern()) - want) <= rtol * max(abs(want), 1.0)


def test_TestBaselines_on_gpu(ctx):
    """gtest_all.cpp:373-419 through the facade with the DEVICE sort (tie order differs from std::sort,
    the partition into sets does not)."""
    g = load_golden("testbaselines_gauss_rp_f32")
    s = host.DPSolver(40, 5, g["a"], g["b"], 0, True, True, dtype=np.float32, ctx=ctx)
    from test_oracle import TESTBASELINES_EXPECTED
    assert [sorted(x.tolist()) for x in s.subsets] == TESTBASELINES_EXPECTED


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("dist,rp", [(0, False), (0, True), (1, False), (1, True), (2, False)])
@pytest.mark.parametrize("sum_mode", [capi.PS_SUM_RUNNING, capi.PS_SUM_PREFIX])
def test_cuda_matches_oracle_seeded(oracle, ctx, dt, dist, rp, sum_mode):
    rs = np.random.RandomState(100 + dist * 7 + rp)
    shapes = [(1, 1), (2, 1), (2, 2), (3, 2), (3, 5), (17, 4), (128, 3), (129, 9), (257, 257), (700, 6), (1500, 12)]
    for n, T in shapes:
        for kind in ("counts", "real"):
            if kind == "counts":
                b = np.maximum(rs.poisson(50, n), 1).astype(dt)
                a = np.maximum(rs.poisson(b.astype(np.float64) * rs.uniform(0.3, 1.7)), 1).astype(dt)
            else:
                b = rs.uniform(0.5, 10, n).astype(dt)
                a = (rs.uniform(0.5, 10, n) if dist == 1 else rs.uniform(-10, 10, n)).astype(dt)
            ref = oracle.solve("port", a, b, T, dist, rp, dtype=dt)
            for chunk in (0, 128, 384):
                _check_against(oracle, ctx, a, b, T, dist, rp, dt, ref, sum_mode=sum_mode, chunk_len=chunk,
                               exact_sums=(kind == "counts"))


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_prefix_mode_is_bit_identical_on_exact_sums(ctx, dt):
    a, b = synth.poisson_planted(5000, 8, seed=5, dtype=dt)
    for dist, rp in ((1, False), (2, False), (0, True)):
        r1 = ctx.solve_raw(a, b, 8, dist, rp, dtype=dt, sum_mode=capi.PS_SUM_RUNNING, levels=True)
        r2 = ctx.solve_raw(a, b, 8, dist, rp, dtype=dt, sum_mode=capi.PS_SUM_PREFIX, levels=True)
        for k in ("perm", "bounds", "subset_scores", "max_score", "next_start"):
            assert np.array_equal(r1[k], r2[k]), k


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_device_sort(oracle, ctx, dt):
    # tie-free keys: the radix sort must reproduce std::sort's permutation exactly
    a, b = synth.gaussian_mixture(4000, 5, seed=3, dtype=dt)
    ref = oracle.solve("port", a, b, 3, 2, False, dtype=dt, levels=False)
    raw = ctx.solve_raw(a, b, 3, 2, False, dtype=dt)
    assert np.array_equal(raw["perm"], ref["perm"])
    # negative keys, massive ties: stable order == argsort(kind='stable') of the same quotients
    rs = np.random.RandomState(4)
    a = rs.randint(-3, 4, 3000).astype(dt)
    b = rs.randint(1, 4, 3000).astype(dt)
    raw = ctx.solve_raw(a, b, 3, 2, False, dtype=dt)
    q = (a / b).astype(dt)
    q = np.where(q == 0, dt(0), q)  # -0.0 == +0.0 under '<'
    want = np.argsort(q, kind="stable")
    assert np.array_equal(np.sort(raw["perm"]), np.arange(3000))
    assert np.all(np.diff(q[raw["perm"]]) >= 0)
    nz = q[raw["perm"]] != 0
    assert np.array_equal(raw["perm"][nz], want[q[want] != 0])


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_batch_equals_single_instances(ctx, dt):
    a, b = synth.poisson_null(777, seed=9, dtype=dt, replicas=9)
    for dist, rp, sweep in ((1, False, False), (1, True, True), (2, False, False)):
        rb = ctx.solve_raw(a, b, 6, dist, rp, dtype=dt, sweep=sweep, levels=True)
        for r in range(9):
            r1 = ctx.solve_raw(a[r], b[r], 6, dist, rp, dtype=dt, sweep=sweep, levels=True)
            for k in ("perm", "bounds", "subset_scores", "max_score", "next_start"):
                assert np.array_equal(rb[k][r], r1[k]), (k, r)


@pytest.mark.parametrize("dist", [0, 1, 2])
def test_TestSweepResultsMatchSingleTResults_on_gpu(ctx, dist):  # gtest_all.cpp:603-662
    rs = np.random.RandomState(21 + dist)
    n, T = 300, 9
    a = rs.uniform(0.1, 10, n).astype(np.float32)
    b = rs.uniform(0.1, 10, n).astype(np.float32)
    sw = host.DPSolver(n, T, a, b, dist, True, True, 0.0, 1, True, False, dtype=np.float32, ctx=ctx)
    allp = sw.get_all_subsets_and_scores_extern()
    assert len(allp) == T + 1 and allp[0][0] == []
    for S in range(1, T + 1):
        one = host.DPSolver(n, S, a, b, dist, True, True, dtype=np.float32, ctx=ctx)
        assert one.get_optimal_score_extern() == allp[S][1]
        assert one.get_optimal_subsets_extern() == allp[S][0]


def test_ols_on_gpu_recovers_planted_levels(ctx):  # solver_tests.py:181-201
    rs = np.random.RandomState(0xC0FFEE3 % (2 ** 31))
    n = 2000
    b = rs.uniform(90, 110, n)
    a = rs.normal(np.repeat([0.25, 0.75, 1.25, 1.75], n // 4) * b, 1.0)
    s = host.DPSolver(n, 12, a, b, 2, True, True, 0.0, 1, False, True, dtype=np.float32, ctx=ctx)
    assert s.get_optimal_num_clusters_OLS_extern() == 4
    assert len(s.get_optimal_subsets_extern()) == 4


def test_range_score(oracle, ctx):  # compute_score, python_dpsolver.cpp:74-111
    rs = np.random.RandomState(2)
    for dt in (np.float32, np.float64):
        a = rs.uniform(0.5, 10, 999).astype(dt)
        b = rs.uniform(0.5, 10, 999).astype(dt)
        for dist in (0, 2):
            for rp in (False, True):
                assert ctx.range_score(a, b, dist, rp, dtype=dt) == oracle.range_score(a, b, dist, rp, 0, 999, dtype=dt)
        got = float(ctx.range_score(a, b, 1, False, dtype=dt))
        want = float(oracle.range_score(a, b, 1, False, 0, 999, dtype=dt))
        assert abs(got - want) <= tol_for(dt, True) * abs(want)


def test_errors(ctx):
    a = np.ones(10, np.float32)
    with pytest.raises(capi.DistributionError):  # distributionException, DP_impl.hpp:62-64
        ctx.solve_raw(a, a, 2, 7, False)
    with pytest.raises(capi.PsError):
        ctx.solve_raw(a, a, 0, 1, False)
    with pytest.raises(capi.PsError):
        ctx.solve_raw(np.ones(0, np.float32), np.ones(0, np.float32), 2, 1, False)


def test_nan_cells_never_win(oracle, ctx):
    # Poisson risk-part with a zero count: S = 0*log(0) = NaN for ranges summing to 0
    rs = np.random.RandomState(8)
    n = 200
    b = np.maximum(rs.poisson(20, n), 1).astype(np.float64)
    a = np.maximum(rs.poisson(b), 1).astype(np.float64)
    a[rs.randint(0, n)] = 0.0
    ref = oracle.solve("port", a, b, 4, 1, True, dtype=np.float64)
    raw = ctx.solve_raw(a, b, 4, 1, True, dtype=np.float64, perm=ref["perm"], levels=True)
    assert np.array_equal(np.isnan(raw["max_score"]), np.isnan(ref["max_score"]))
    assert np.array_equal(raw["next_start"], ref["next_start"])


# ------------------------------------------------------------------ full BASELINE sizes
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_full_size_C3_properties_and_sampled_rows(oracle, ctx, dt):
    cfg = synth.CONFIGS["C3"]
    n, T = cfg["n"], cfg["T"]
    a, b = cfg["gen"](dt)
    r1 = ctx.solve_raw(a, b, T, 1, False, dtype=dt, sum_mode=capi.PS_SUM_RUNNING, levels=True)
    # (1) counts: PREFIX and RUNNING are bit-identical, and the tiling does not change anything
    r2 = ctx.solve_raw(a, b, T, 1, False, dtype=dt, sum_mode=capi.PS_SUM_PREFIX, levels=True, chunk_len=8192)
    for k in ("perm", "bounds", "subset_scores", "max_score", "next_start"):
        assert np.array_equal(r1[k], r2[k]), k
    # (2) structure: strictly increasing boundaries ending at n; the stable sort is a permutation
    bd = r1["bounds"][0]
    assert bd[0] == 0 and bd[-1] == n and np.all(np.diff(bd) > 0)
    assert np.array_equal(np.sort(r1["perm"]), np.arange(n))
    # (3) the objective is the sum of the subset scores the backtrace reports
    tot = float(np.sum(r1["subset_scores"][0].astype(np.float64)))
    assert abs(float(r1["max_score"][T, 0]) - tot) <= (1e-12 if dt == np.float64 else 1e-5) * abs(tot)
    # (4) sampled rows of three levels against the oracle, fed with the GPU's own previous level
    perm = r1["perm"]
    a_s, b_s = a[perm], b[perm]
    rs = np.random.RandomState(1)
    rtol, tie = tol_for(dt, True), TIE[dt]
    exact = rtol == 0.0
    for j in (2, 9, T):
        kmax = n - j + 1
        rows = np.array([0], np.int32) if j == T else np.unique(
            np.concatenate([[0, 1, 127, 128, 4095, 4096, 4097, kmax - 1, kmax - 2],
                            rs.randint(0, kmax, 24)])).astype(np.int32)
        best, arg = oracle.dp_rows(a_s, b_s, r1["max_score"][j - 1], kmax, 1, False, rows, dtype=dt)
        got_v, got_k = r1["max_score"][j][rows], r1["next_start"][j][rows]
        scale = np.max(np.abs(r1["max_score"][j][:kmax if j < T else 1].astype(np.float64)))
        if exact:
            assert np.array_equal(got_v, best) and np.array_equal(got_k, arg), f"level {j}: sampled rows differ"
        assert np.all(np.abs(got_v.astype(np.float64) - best.astype(np.float64)) <= rtol * scale)
        for q in np.nonzero(got_k != arg)[0]:
            i = int(rows[q])
            vg = float(oracle.range_score(a_s, b_s, 1, False, i, int(got_k[q]), dtype=dt)) + float(
                r1["max_score"][j - 1][got_k[q]])
            vo = float(oracle.range_score(a_s, b_s, 1, False, i, int(arg[q]), dtype=dt)) + float(
                r1["max_score"][j - 1][arg[q]])
            assert abs(vg - vo) <= tie * max(abs(vo), scale)


def test_full_size_C4_replica_shape(oracle, ctx):
    """C4 shape at a reduced replica count: batched RP solve == oracle per replica (f32 Poisson)."""
    n, T, R = 5000, 8, 6
    a, b = synth.poisson_null(n, seed=synth.SEED + 4, dtype=np.float32, replicas=R)
    refs = [oracle.solve("port", a[r], b[r], T, 1, True, dtype=np.float32) for r in range(R)]
    perm = np.stack([x["perm"] for x in refs])
    raw = ctx.solve_raw(a, b, T, 1, True, dtype=np.float32, perm=perm, levels=True)
    for r in range(R):
        a_s, b_s = a[r][perm[r]], b[r][perm[r]]
        check_levels(raw["max_score"][r], raw["next_start"][r], refs[r]["max_score"], refs[r]["next_start"],
                     tol_for(np.float32, True), TIE[np.float32],
                     _recompute_fn(oracle, a_s, b_s, refs[r]["max_score"], 1, True, np.float32))


def test_device_pointer_call_equals_host_buffer_call(ctx):
    """ps_problem.on_device = 1 (inputs and outputs resident in HBM, what bench.py's `value` times) gives the
    same boundaries, scores and sort index as the host-buffer call."""
    import torch
    dev = torch.device("cuda", ctx.device)
    a, b = synth.poisson_planted(6000, 9, seed=71, dtype=np.float64)
    want = ctx.solve_raw(a, b, 9, 1, False, dtype=np.float64, sweep=True)
    a_d, b_d = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
    perm_d = torch.empty(6000, dtype=torch.int32, device=dev)
    bounds_d = torch.empty((10, 10), dtype=torch.int32, device=dev)
    sc_d = torch.empty((10, 9), dtype=torch.float64, device=dev)
    ctx.solve_device(a_d.data_ptr(), b_d.data_ptr(), 6000, 9, 1, False, np.float64, sweep=True,
                     d_perm_out=perm_d.data_ptr(), d_bounds=bounds_d.data_ptr(), d_scores=sc_d.data_ptr())
    assert np.array_equal(perm_d.cpu().numpy(), want["perm"])
    assert np.array_equal(bounds_d.cpu().numpy(), want["bounds"])
    assert np.array_equal(sc_d.cpu().numpy(), want["subset_scores"])


def test_batched_sweep_prefix_mode_and_limits(oracle, ctx):
    a, b = synth.poisson_null(400, seed=72, dtype=np.float32, replicas=5)
    rb = ctx.solve_raw(a, b, 6, 1, True, dtype=np.float32, sweep=True, sum_mode=capi.PS_SUM_PREFIX)
    for r in range(5):
        ref = oracle.solve("port", a[r], b[r], 6, 1, True, dtype=np.float32, sweep=True)
        if np.array_equal(rb["perm"][r], ref["perm"]):  # tie order may differ (stable vs std::sort)
            s = host.DPSolver(400, 6, a[r], b[r], 1, True, True, 0.0, 1, True, False, dtype=np.float32, ctx=ctx,
                              sum_mode=capi.PS_SUM_PREFIX)
            for S in range(1, 7):
                assert s.all[S][1] == ref["all"][S][1]
    # more than 65535 instances (the instance index is a grid dimension) run as slices: see
    # test_batches_above_the_grid_limit_are_sliced
    big = ctx.solve_raw(np.ones((65536, 4), np.float32), np.ones((65536, 4), np.float32), 2, 2, False)
    assert big["bounds"].shape == (65536, 1, 3) and np.array_equal(big["bounds"][65535], big["bounds"][0])


def test_one_million_elements_sampled_rows(oracle, ctx):
    """BASELINE configs[4] size on one GPU (n = 1,000,000; T = 3 keeps it to one full level): sampled rows of level 2
    and the final row, recomputed by the oracle from the GPU's own level-1 row, must agree bit for bit -- guards the
    64-bit index arithmetic, the 16384-wide chunks and the 62-chunk SA/SB tables."""
    n, T = 1_000_000, 3
    a, b = synth.poisson_planted(n, 8, seed=81, dtype=np.float64)
    r = ctx.solve_raw(a, b, T, 1, False, dtype=np.float64, levels=True)
    a_s, b_s = a[r["perm"]], b[r["perm"]]
    bd = r["bounds"][0]
    assert bd[0] == 0 and bd[-1] == n and np.all(np.diff(bd) > 0)
    rows = np.array([0, 1, 127, 128, 16383, 16384, 16385, 500_000, 999_000, n - 3, n - 2], np.int32)
    best, arg = oracle.dp_rows(a_s, b_s, r["max_score"][1], n - 1, 1, False, rows, dtype=np.float64)
    assert np.array_equal(r["max_score"][2][rows], best)
    assert np.array_equal(r["next_start"][2][rows], arg)
    best, arg = oracle.dp_rows(a_s, b_s, r["max_score"][2], n - 2, 1, False, np.array([0], np.int32), dtype=np.float64)
    assert r["max_score"][3][0] == best[0] and r["next_start"][3][0] == arg[0]
    # level 1 itself: S(i, n) for a few i
    for i in (0, 12345, n - 1):
        assert r["max_score"][1][i] == oracle.range_score(a_s, b_s, 1, False, i, n, dtype=np.float64)


def test_large_host_batch_is_pipelined_with_identical_results(ctx):
    """Host-buffer batches of >= 256 instances and >= 16 MB are cut into parts whose copies overlap the kernels of the
    parts before them (solve_pipelined): same outputs as the same instances solved in smaller, unpipelined calls."""
    n, T, R = 7000, 4, 320
    a, b = synth.poisson_null(n, seed=synth.SEED + 90, dtype=np.float64, replicas=R)
    whole = ctx.solve_raw(a, b, T, 1, False, dtype=np.float64, levels=True, sweep=True)
    t = ctx.timings()
    assert t["h2d_bytes"] == 2 * R * n * 8 and t["kernel_launches"] > 0
    for lo, hi in ((0, 150), (150, 320)):
        part = ctx.solve_raw(a[lo:hi], b[lo:hi], T, 1, False, dtype=np.float64, levels=True, sweep=True)
        for k in ("perm", "bounds", "subset_scores", "max_score", "next_start"):
            assert np.array_equal(whole[k][lo:hi], part[k]), (k, lo)


def test_batches_above_the_grid_limit_are_sliced(ctx):
    """ps_problem.batch > 65535 (the instance index is a grid dimension): solved as consecutive slices, same results as
    instance-by-instance calls -- checked on the instances around the slice boundary and at both ends."""
    R, n, T = 65535 + 70, 12, 3
    rs = np.random.RandomState(3)
    a = rs.uniform(0.5, 9.5, (R, n)).astype(np.float32)
    b = rs.uniform(0.5, 9.5, (R, n)).astype(np.float32)
    raw = ctx.solve_raw(a, b, T, 2, False, dtype=np.float32)
    for r in (0, 1, 65533, 65534, 65535, 65536, R - 1):
        one = ctx.solve_raw(a[r], b[r], T, 2, False, dtype=np.float32)
        assert np.array_equal(raw["perm"][r], one["perm"]) and np.array_equal(raw["bounds"][r], one["bounds"])
        assert np.array_equal(raw["subset_scores"][r], one["subset_scores"]), r
