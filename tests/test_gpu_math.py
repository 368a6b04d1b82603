"""GPU tests of the device math the level kernels are built on: branch-free IEEE division (csrc/math_fast.cuh)
must equal div.rn bit for bit on the safe range; the log/logf restatements of glibc (csrc/glibc_log.cuh) must
equal the HOST's libm bit for bit (that is what makes the Poisson path bit-identical to the reference); and the
level kernels on the fast path must produce exactly the tables of the general kernels."""
import numpy as np
import pytest

from partitionsolvers_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("which,name", [(0, "fdiv"), (1, "logf entries"), (2, "ddiv"), (3, "log entries")])
def test_fast_math_is_bit_identical_to_general(ctx, which, name):
    assert ctx.selftest_math(which, 1 << 30, seed=which + 1) == 0, name


def _log_args(dtype, n=400000):
    rs = np.random.RandomState(12)
    near1 = 1.0 + (rs.uniform(-0.07, 0.07, n))
    ratio = rs.uniform(0.05, 25.0, n)
    wide = np.ldexp(rs.uniform(1.0, 2.0, n), rs.randint(-120 if dtype == np.float32 else -1000,
                                                          120 if dtype == np.float32 else 1000, n))
    tiny = np.finfo(dtype).tiny
    special = np.array([0.0, -0.0, 1.0, np.inf, -np.inf, np.nan, -1.0, tiny, tiny / 4, tiny / 1024,
                        np.finfo(dtype).max, 0.9375, 1.064697265625, np.nextafter(dtype(1), dtype(0)),
                        np.nextafter(dtype(1), dtype(2))], dtype=np.float64)
    return np.concatenate([near1, ratio, wide, special]).astype(dtype)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_device_log_equals_host_libm_bit_for_bit(oracle, ctx, dt):
    x = _log_args(dt)
    with np.errstate(all="ignore"):
        want = oracle.host_log(x, dtype=dt)
    got_general, got_fast = ctx.eval_log(x, dtype=dt)
    it = np.uint64 if dt == np.float64 else np.uint32
    nan = np.isnan(want)
    assert np.array_equal(np.isnan(got_general), nan)
    same = got_general.view(it) == want.view(it)
    assert np.all(same | nan), f"{int((~(same | nan)).sum())} arguments differ from the host libm"
    normal = np.isfinite(x) & (x >= np.finfo(dt).tiny)
    assert np.array_equal(got_fast[normal].view(it), want[normal].view(it))


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("dist,rp", [(1, False), (1, True), (0, False), (0, True), (2, False)])
def test_fast_and_generic_level_kernels_agree_bitwise(ctx, dt, dist, rp):
    for gen in (lambda: synth.poisson_planted(3000, 7, seed=61, dtype=dt),
                lambda: synth.uniform_ab(2000, seed=62, dtype=dt, a_lo=0.01, b_lo=0.01)):
        a, b = gen()
        r_fast = ctx.solve_raw(a, b, 7, dist, rp, dtype=dt, levels=True)
        r_gen = ctx.solve_raw(a, b, 7, dist, rp, dtype=dt, levels=True, no_fast=1)
        r_scalar = ctx.solve_raw(a, b, 7, dist, rp, dtype=dt, levels=True, no_fast=2, chunk_len=512)
        for k in ("perm", "bounds", "subset_scores", "max_score", "next_start"):
            assert np.array_equal(r_fast[k], r_gen[k], equal_nan=True), k
            assert np.array_equal(r_fast[k], r_scalar[k], equal_nan=True), k


def test_out_of_range_inputs_fall_back_to_generic_path(ctx, oracle):
    # a tiny b and a huge b (outside the prechecked range): results must still match the oracle bit for bit
    rs = np.random.RandomState(3)
    n = 600
    b = rs.uniform(0.5, 10, n)
    a = rs.uniform(0.5, 10, n)
    b[5] = 1e-12
    b[17] = 1e12
    for dist in (2, 1):
        ref = oracle.solve("port", a, b, 4, dist, False, dtype=np.float64)
        raw = ctx.solve_raw(a, b, 4, dist, False, dtype=np.float64, perm=ref["perm"], levels=True)
        assert np.array_equal(raw["max_score"], ref["max_score"])
        assert np.array_equal(raw["next_start"], ref["next_start"])
