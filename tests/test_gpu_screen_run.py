"""GPU tests of the running-sum screened kernel (csrc/screen_run.cuh): double inputs whose partial sums round (real-valued
a, b).  Bounds come from fixed-point prefix sums, surviving cells from the reference's running sums; the tables must be
those of the exhaustive kernel and of the oracle, bit for bit, and the audit must find no bound that was too low."""
import numpy as np
import pytest

from partitionsolvers_b200 import capi, synth

pytestmark = pytest.mark.gpu

SCORES = [(1, False), (1, True), (0, False), (0, True), (2, False)]
F64 = np.float64


def _same_tables(r0, r1):
    for k in ("max_score", "next_start", "bounds", "subset_scores", "perm"):
        assert np.array_equal(r0[k], r1[k]), k


def _real_valued(n, levels, seed, positive):
    a, b = synth.gaussian_mixture(n, levels, seed=seed, dtype=F64)
    if positive:
        a = np.abs(a) + 0.25
    return a, b


@pytest.mark.parametrize("dist,rp", SCORES)
def test_running_screen_equals_exhaustive_and_audit_is_clean(ctx, dist, rp):
    n, T = 20000, 6
    a, b = _real_valued(n, 6, synth.SEED + 60, positive=(dist == 1))
    exh = ctx.solve_raw(a, b, T, dist, rp, dtype=F64, levels=True, sweep=True, screen=1)
    assert ctx.timings()["screen_used"] == 0
    scr = ctx.solve_raw(a, b, T, dist, rp, dtype=F64, levels=True, sweep=True, screen=0)
    t = ctx.timings()
    assert t["screen_used"] == 2
    cells = sum(t["level_cells"][:-1])
    assert 0 < t["screen_exact_cells"] < 0.5 * cells
    _same_tables(exh, scr)
    aud = ctx.solve_raw(a, b, T, dist, rp, dtype=F64, levels=True, sweep=True, screen=2)
    ta = ctx.timings()
    assert ta["screen_used"] == 2 and ta["screen_violations"] == 0
    _same_tables(exh, aud)


def test_running_screen_matches_oracle(oracle, ctx):
    """Against the CPU oracle: every table cell identical (audit mode takes the running-sum screen at any size)."""
    n, T = 6100, 5
    for name, (a, b) in (("tract", synth.tract_like(n, 4, seed=synth.SEED + 61, dtype=F64)),
                         ("mixture", _real_valued(n, 5, synth.SEED + 62, positive=False))):
        for dist, rp in SCORES:
            aa = np.maximum(a, 0.5) if dist == 1 else a
            ref = oracle.solve("port", aa, b, T, dist, rp, dtype=F64)
            raw = ctx.solve_raw(aa, b, T, dist, rp, dtype=F64, perm=ref["perm"], levels=True, screen=2)
            t = ctx.timings()
            assert t["screen_used"] == 2 and t["screen_violations"] == 0, (name, dist, rp)
            assert np.array_equal(raw["next_start"], ref["next_start"]), (name, dist, rp)
            assert np.array_equal(raw["max_score"], ref["max_score"]), (name, dist, rp)


def test_rows_with_nonpositive_sums_are_evaluated_exhaustively(ctx):
    """Elements with a <= 0 sort to the front; the row blocks that contain them take no bounds (condition F4)."""
    n, T = 17000, 5
    a, b = _real_valued(n, 8, synth.SEED + 63, positive=False)
    assert 0 < (a <= 0).sum() <= n // 8
    for dist, rp in ((0, False), (0, True), (2, False)):
        exh = ctx.solve_raw(a, b, T, dist, rp, dtype=F64, levels=True, screen=1)
        scr = ctx.solve_raw(a, b, T, dist, rp, dtype=F64, levels=True, screen=0)
        assert ctx.timings()["screen_used"] == 2
        _same_tables(exh, scr)


def test_unsorted_caller_permutation(ctx):
    n, T = 6500, 5
    a, b = _real_valued(n, 5, synth.SEED + 64, positive=True)
    perm = np.random.RandomState(6).permutation(n).astype(np.int32)
    for dist, rp in SCORES:
        exh = ctx.solve_raw(a, b, T, dist, rp, dtype=F64, perm=perm, levels=True, screen=1)
        aud = ctx.solve_raw(a, b, T, dist, rp, dtype=F64, perm=perm, levels=True, screen=2)
        t = ctx.timings()
        assert t["screen_used"] == 2 and t["screen_violations"] == 0, (dist, rp)
        _same_tables(exh, aud)


def test_batch_of_instances(ctx):
    n, T, R = 16500, 4, 3
    ab = [_real_valued(n, 4 + r, synth.SEED + 65 + r, positive=True) for r in range(R)]
    a = np.stack([x[0] for x in ab])
    b = np.stack([x[1] for x in ab])
    exh = ctx.solve_raw(a, b, T, 2, False, dtype=F64, levels=True, screen=1)
    scr = ctx.solve_raw(a, b, T, 2, False, dtype=F64, levels=True, screen=0)
    assert ctx.timings()["screen_used"] == 2
    _same_tables(exh, scr)


def test_conditions_of_the_fixed_point_bounds(ctx):
    n, T = 17000, 4
    a, b = _real_valued(n, 4, synth.SEED + 68, positive=True)
    # (F1) one b twelve orders of magnitude below the rest: its fixed-point image would lose all relative accuracy
    b2 = b.copy()
    b2[5] = 1e-8
    ctx.solve_raw(a, b2, T, 2, False, dtype=F64)
    assert ctx.timings()["screen_used"] == 0
    # (F4) too many non-positive a
    a2 = a.copy()
    a2[: n // 4] *= -1.0
    ctx.solve_raw(a2, b, T, 2, False, dtype=F64)
    assert ctx.timings()["screen_used"] == 0
    # (F2) Poisson risk-part needs the two-sided condition on a as well; the other scores only bound a from above
    a3 = a.copy()
    a3[7] = 1e-8
    ctx.solve_raw(a3, b, T, 1, True, dtype=F64)
    assert ctx.timings()["screen_used"] == 0
    exh = ctx.solve_raw(a3, b, T, 1, False, dtype=F64, levels=True, screen=1)
    scr = ctx.solve_raw(a3, b, T, 1, False, dtype=F64, levels=True)
    assert ctx.timings()["screen_used"] == 2
    _same_tables(exh, scr)
    # float inputs: Poisson risk-part is screened too (box bound on the running sums themselves) and gives the same tables
    af, bf = a.astype(np.float32), b.astype(np.float32)
    exh = ctx.solve_raw(af, bf, T, 1, True, dtype=np.float32, levels=True, screen=1)
    scr = ctx.solve_raw(af, bf, T, 1, True, dtype=np.float32, levels=True)
    assert ctx.timings()["screen_used"] == 2
    _same_tables(exh, scr)


def test_full_size_rational(ctx):
    """BASELINE configuration 3, Rational score on real-valued data (n = 100000), T = 4: identical tables."""
    a, b = synth.CONFIGS["C3R"]["gen"](F64)
    exh = ctx.solve_raw(a, b, 4, 2, False, dtype=F64, levels=True, screen=1)
    scr = ctx.solve_raw(a, b, 4, 2, False, dtype=F64, levels=True, screen=0)
    assert ctx.timings()["screen_used"] == 2
    _same_tables(exh, scr)


# ------------------------------------------------------------------------------------------------ float inputs
F32 = np.float32
F32_SCORES = [(0, False), (0, True), (2, False), (1, False), (1, True)]  # every score (screen_run.cuh)


@pytest.mark.parametrize("dist,rp", F32_SCORES)
def test_float_running_screen_equals_exhaustive_and_audit_is_clean(ctx, dist, rp):
    n, T = 20000, 6
    for positive in ((True,) if dist == 1 else (True, False)):  # False: some a <= 0, the first row blocks take no bounds
        a, b = _real_valued(n, 6, synth.SEED + 70, positive=positive)
        a, b = a.astype(F32), b.astype(F32)
        exh = ctx.solve_raw(a, b, T, dist, rp, dtype=F32, levels=True, sweep=True, screen=1)
        assert ctx.timings()["screen_used"] == 0
        scr = ctx.solve_raw(a, b, T, dist, rp, dtype=F32, levels=True, sweep=True, screen=0)
        t = ctx.timings()
        assert t["screen_used"] == 2
        assert 0 < t["screen_exact_cells"] < 0.5 * sum(t["level_cells"][:-1])
        _same_tables(exh, scr)
        aud = ctx.solve_raw(a, b, T, dist, rp, dtype=F32, levels=True, sweep=True, screen=2)
        ta = ctx.timings()
        assert ta["screen_used"] == 2 and ta["screen_violations"] == 0
        _same_tables(exh, aud)


def test_float_running_screen_matches_oracle(oracle, ctx):
    n, T = 6100, 5
    for name, (a, b) in (("tract", synth.tract_like(n, 4, seed=synth.SEED + 71, dtype=F32)),
                         ("mixture", synth.gaussian_mixture(n, 5, seed=synth.SEED + 72, dtype=F32))):
        for dist, rp in F32_SCORES:
            if dist == 1:
                a = np.maximum(a, F32(0.5))  # Poisson needs positive a
            ref = oracle.solve("port", a, b, T, dist, rp, dtype=F32)
            raw = ctx.solve_raw(a, b, T, dist, rp, dtype=F32, perm=ref["perm"], levels=True, screen=2)
            t = ctx.timings()
            assert t["screen_used"] == 2 and t["screen_violations"] == 0, (name, dist, rp)
            assert np.array_equal(raw["next_start"], ref["next_start"]), (name, dist, rp)
            assert np.array_equal(raw["max_score"], ref["max_score"]), (name, dist, rp)


def test_float_running_screen_unsorted_permutation_and_batch(ctx):
    n, T = 6500, 5
    a, b = synth.gaussian_mixture(n, 5, seed=synth.SEED + 73, dtype=F32)
    perm = np.random.RandomState(7).permutation(n).astype(np.int32)
    for dist, rp in F32_SCORES:
        if dist == 1:
            a = np.maximum(a, F32(0.5))
        exh = ctx.solve_raw(a, b, T, dist, rp, dtype=F32, perm=perm, levels=True, screen=1)
        aud = ctx.solve_raw(a, b, T, dist, rp, dtype=F32, perm=perm, levels=True, screen=2)
        t = ctx.timings()
        assert t["screen_used"] == 2 and t["screen_violations"] == 0, (dist, rp)
        _same_tables(exh, aud)
    n, R = 16500, 3
    ab = [synth.gaussian_mixture(n, 4 + r, seed=synth.SEED + 74 + r, dtype=F32) for r in range(R)]
    a = np.stack([x[0] for x in ab])
    b = np.stack([x[1] for x in ab])
    exh = ctx.solve_raw(a, b, 4, 2, False, dtype=F32, levels=True, sweep=True, screen=1)
    scr = ctx.solve_raw(a, b, 4, 2, False, dtype=F32, levels=True, sweep=True, screen=0)
    assert ctx.timings()["screen_used"] == 2
    _same_tables(exh, scr)


def test_full_size_rational_float(ctx):
    a, b = synth.CONFIGS["C3R"]["gen"](F32)
    exh = ctx.solve_raw(a, b, 4, 2, False, dtype=F32, levels=True, screen=1)
    scr = ctx.solve_raw(a, b, 4, 2, False, dtype=F32, levels=True, screen=0)
    assert ctx.timings()["screen_used"] == 2
    _same_tables(exh, scr)


@pytest.mark.parametrize("rp", [False, True])
def test_full_size_poisson_float_real_valued(ctx, rp):
    """BASELINE configuration 3 shape, Poisson (both objectives) on REAL-VALUED float data (sums round): the float
    running-sum screen, audited at T = 3 and plain at T = 4, against the exhaustive kernel."""
    a, b = synth.CONFIGS["C3R"]["gen"](F32)
    a = np.abs(a) + F32(0.25)
    exh = ctx.solve_raw(a, b, 4, 1, rp, dtype=F32, levels=True, screen=1)
    assert ctx.timings()["screen_used"] == 0
    scr = ctx.solve_raw(a, b, 4, 1, rp, dtype=F32, levels=True, screen=0)
    assert ctx.timings()["screen_used"] == 2
    _same_tables(exh, scr)
    aud = ctx.solve_raw(a, b, 3, 1, rp, dtype=F32, levels=True, screen=2)
    t = ctx.timings()
    assert t["screen_used"] == 2 and t["screen_violations"] == 0
    # level 3 of the T = 3 solve is its last level: row 0 only
    assert np.array_equal(aud["max_score"][:3], exh["max_score"][:3]) and aud["max_score"][3][0] == exh["max_score"][3][0]
    assert np.array_equal(aud["next_start"][:3], exh["next_start"][:3]) and aud["next_start"][3][0] == exh["next_start"][3][0]
