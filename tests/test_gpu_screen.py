"""GPU tests of the screened level kernel (csrc/screen.cuh): it must produce the SAME tables as the exhaustive kernel
and as the oracle, bit for bit; its bounds are audited cell by cell; instances that do not satisfy its preconditions
must take the exhaustive kernel."""
import numpy as np
import pytest

from partitionsolvers_b200 import capi, synth

pytestmark = pytest.mark.gpu

SCORES = [(1, False), (1, True), (0, False), (0, True), (2, False)]


def _same_tables(r0, r1):
    for k in ("max_score", "next_start", "bounds", "subset_scores", "perm"):
        assert np.array_equal(r0[k], r1[k]), k


def test_mufu_error_claims(ctx):
    """The error budget of the float bounds assumes |x rcp(x) - 1| <= 4u and |lg2 x - log2 x| <= 4u max(1, |log2 x|)
    (u = 2^-24); checked here over EVERY float of the safe range."""
    rcp, lg2 = ctx.selftest_mufu()
    assert 0 < rcp <= 4.0, rcp
    assert 0 < lg2 <= 4.0, lg2


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("dist,rp", SCORES)
def test_screened_equals_exhaustive_and_audit_is_clean(ctx, dt, dist, rp):
    n, T = 6000, 6
    a, b = synth.poisson_planted(n, 6, seed=synth.SEED + 40, dtype=dt)
    exh = ctx.solve_raw(a, b, T, dist, rp, dtype=dt, levels=True, sweep=True, screen=1)
    assert ctx.timings()["screen_used"] == 0
    scr = ctx.solve_raw(a, b, T, dist, rp, dtype=dt, levels=True, sweep=True, screen=0)
    t = ctx.timings()
    assert t["screen_used"] == 1
    cells = sum(t["level_cells"][:-1])
    assert 0 < t["screen_exact_cells"] < 0.5 * cells  # it does screen
    _same_tables(exh, scr)
    aud = ctx.solve_raw(a, b, T, dist, rp, dtype=dt, levels=True, sweep=True, screen=2)
    ta = ctx.timings()
    assert ta["screen_used"] == 1 and ta["screen_violations"] == 0
    _same_tables(exh, aud)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_screened_matches_oracle(oracle, ctx, dt):
    """Against the CPU oracle (running sums, the reference's order): every table cell identical."""
    n, T = 6100, 5
    a, b = synth.poisson_planted(n, 5, seed=synth.SEED + 41, dtype=dt)
    for dist, rp in SCORES:
        ref = oracle.solve("port", a, b, T, dist, rp, dtype=dt)
        raw = ctx.solve_raw(a, b, T, dist, rp, dtype=dt, perm=ref["perm"], levels=True)
        assert ctx.timings()["screen_used"] == 1
        assert np.array_equal(raw["next_start"], ref["next_start"]), (dist, rp)
        assert np.array_equal(raw["max_score"], ref["max_score"]), (dist, rp)


def test_null_model_batch_and_dyadic_inputs(ctx):
    """Replica batches (flat objective: the worst case for screening) and dyadic non-integer inputs."""
    a, b = synth.poisson_null(5000, seed=synth.SEED + 42, dtype=np.float32, replicas=16)
    exh = ctx.solve_raw(a, b, 8, 1, False, dtype=np.float32, levels=True, screen=1)
    aud = ctx.solve_raw(a, b, 8, 1, False, dtype=np.float32, levels=True, screen=2)
    t = ctx.timings()
    assert t["screen_used"] == 1 and t["screen_violations"] == 0
    _same_tables(exh, aud)
    a, b = synth.tract_like(6500, 4, seed=synth.SEED + 43, dtype=np.float64, dyadic_bits=10)
    a = np.maximum(a, 1.0)
    for dist, rp in SCORES:
        exh = ctx.solve_raw(a, b, 5, dist, rp, dtype=np.float64, levels=True, screen=1)
        aud = ctx.solve_raw(a, b, 5, dist, rp, dtype=np.float64, levels=True, screen=2)
        t = ctx.timings()
        assert t["screen_used"] == 1 and t["screen_violations"] == 0, (dist, rp)
        _same_tables(exh, aud)


def test_preconditions_send_other_inputs_to_the_exhaustive_kernel(ctx):
    n, T = 7000, 5
    # real-valued b: partial sums are not exact, prefix differences would not be the reference's running sums
    a, b = synth.tract_like(n, 4, seed=synth.SEED + 44, dtype=np.float64)
    ctx.solve_raw(a, b, T, 0, True, dtype=np.float64)
    assert ctx.timings()["screen_used"] == 0
    # a value <= 0 (zero count): the relative-error argument of the float sums needs positive terms
    a, b = synth.poisson_planted(n, 5, seed=synth.SEED + 45, dtype=np.float64)
    a[17] = 0.0
    ctx.solve_raw(a, b, T, 0, False, dtype=np.float64)
    assert ctx.timings()["screen_used"] == 0
    # float sums beyond 2^24 are not exact in float
    a, b = synth.poisson_planted(200000, 5, seed=synth.SEED + 46, dtype=np.float32)
    assert float(a.astype(np.float64).sum()) > 2 ** 24
    # ... so the exact-sum screen is out; the float running-sum screen (screen_run.cuh) takes over, same tables
    scr = ctx.solve_raw(a, b, 3, 1, False, dtype=np.float32, levels=True)
    assert ctx.timings()["screen_used"] == 2
    exh = ctx.solve_raw(a, b, 3, 1, False, dtype=np.float32, levels=True, screen=1)
    assert ctx.timings()["screen_used"] == 0
    _same_tables(exh, scr)
    # ... but fine in double
    ctx.solve_raw(a.astype(np.float64), b.astype(np.float64), 3, 1, False, dtype=np.float64)
    assert ctx.timings()["screen_used"] == 1
    # the user's explicit PREFIX mode and the debug switches keep their kernels
    a, b = synth.poisson_planted(n, 5, seed=synth.SEED + 45, dtype=np.float64)
    ctx.solve_raw(a, b, T, 1, False, dtype=np.float64, sum_mode=capi.PS_SUM_PREFIX)
    assert ctx.timings()["screen_used"] == 0
    ctx.solve_raw(a, b, T, 1, False, dtype=np.float64, no_fast=1)
    assert ctx.timings()["screen_used"] == 0


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_unsorted_caller_permutation_disables_the_monotone_tiers(ctx, dt):
    """With a caller-supplied permutation that is NOT priority-sorted the DP is still well defined (the reference would
    compute it on whatever order it is given); the stage/group tiers rely on ascending priorities, are switched off by
    the device-side sortedness check, and the per-cell tier alone must reproduce the exhaustive result."""
    n, T = 6500, 5
    a, b = synth.poisson_planted(n, 5, seed=synth.SEED + 47, dtype=dt)
    perm = np.random.RandomState(5).permutation(n).astype(np.int32)
    for dist, rp in SCORES:
        exh = ctx.solve_raw(a, b, T, dist, rp, dtype=dt, perm=perm, levels=True, screen=1)
        aud = ctx.solve_raw(a, b, T, dist, rp, dtype=dt, perm=perm, levels=True, screen=2)
        t = ctx.timings()
        assert t["screen_used"] == 1 and t["screen_violations"] == 0, (dist, rp)
        _same_tables(exh, aud)


def test_full_size_screened_equals_exhaustive(ctx):
    """BASELINE configuration 3 (n = 100000), T = 4 to keep the exhaustive side short: identical tables."""
    cfg = synth.CONFIGS["C3"]
    for dt in (np.float64, np.float32):
        a, b = cfg["gen"](dt)
        exh = ctx.solve_raw(a, b, 4, 1, False, dtype=dt, levels=True, screen=1)
        scr = ctx.solve_raw(a, b, 4, 1, False, dtype=dt, levels=True, screen=0)
        assert ctx.timings()["screen_used"] == 1
        _same_tables(exh, scr)
