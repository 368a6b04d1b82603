"""GPU tests of the range-prechecked branch-free math path (csrc/math_fast.cuh): the routines must be
bit-identical to the CUDA math library on the safe range, and the level kernels built on them must produce
exactly the tables of the generic kernels."""
import numpy as np
import pytest

from partitionsolvers_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("which,name", [(0, "fdiv"), (1, "logf"), (2, "ddiv"), (3, "log")])
def test_fast_math_is_bit_identical_to_library(ctx, which, name):
    assert ctx.selftest_math(which, 1 << 30, seed=which + 1) == 0, name


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
    # a zero count (Poisson needs a > 0 for the fast path), a tiny b and a huge b: results must still match the oracle
    rs = np.random.RandomState(3)
    n = 600
    b = rs.uniform(0.5, 10, n)
    a = rs.uniform(0.5, 10, n)
    b[5] = 1e-12
    b[17] = 1e12
    ref = oracle.solve("port", a, b, 4, 2, False, dtype=np.float64)
    raw = ctx.solve_raw(a, b, 4, 2, False, dtype=np.float64, perm=ref["perm"], levels=True)
    assert np.array_equal(raw["max_score"], ref["max_score"])
    assert np.array_equal(raw["next_start"], ref["next_start"])
