"""Seeded synthetic workloads for the consecutive-partitions DP.

Every generator uses numpy's legacy ``RandomState`` stream (stable across numpy versions) so that the
CPU checkers, the golden fixtures and the CUDA path all see identical arrays.  Shapes follow the
reference's own experiments: planted rate levels as in ``gtest_all.cpp:93-162`` /
``SimulatedExperiments.py:139-144``, the null model ``b~Poisson(lam), a~Poisson(b)`` of
``SimulatedExperiments.py:79-80``, and the Gaussian mixture of ``solver_tests.py:145-146``.
"""
import numpy as np

SEED = 0xC0FFEE3

GAUSSIAN, POISSON, RATIONAL = 0, 1, 2


def _levels(n, num_levels, lo=0.1, hi=1.9):
    q = np.linspace(lo, hi, num_levels)
    edges = np.linspace(0, n, num_levels + 1).astype(np.int64)
    out = np.empty(n, np.float64)
    for l in range(num_levels):
        out[edges[l]:edges[l + 1]] = q[l]
    return out


def poisson_planted(n, num_levels, lam=100.0, seed=SEED, dtype=np.float32, shuffle=True):
    """Integer counts: every range sum is exactly representable (f32 while n*lam*1.9 < 2**24)."""
    rs = np.random.RandomState(seed)
    b = rs.poisson(lam, size=n).astype(np.float64)
    b = np.maximum(b, 1.0)
    a = rs.poisson(_levels(n, num_levels) * b).astype(np.float64)
    a = np.maximum(a, 1.0)
    if shuffle:
        p = rs.permutation(n)
        a, b = a[p], b[p]
    return a.astype(dtype), b.astype(dtype)


def poisson_null(n, lam=100.0, seed=SEED, dtype=np.float32, replicas=None):
    """Randomization-test replica(s): b~Poisson(lam), a~Poisson(b)."""
    rs = np.random.RandomState(seed)
    shape = (n,) if replicas is None else (replicas, n)
    b = np.maximum(rs.poisson(lam, size=shape), 1).astype(np.float64)
    a = np.maximum(rs.poisson(b), 1).astype(np.float64)
    return a.astype(dtype), b.astype(dtype)


def tract_like(n, num_levels=4, seed=SEED, dtype=np.float64, dyadic_bits=None):
    """NYC-tract shape: expected counts h=exp(N(3,.5)) (continuous, tie-free), observed g=Poisson(q*h).

    ``dyadic_bits`` rounds h to a 2**-bits grid so that all range sums are exact in f64."""
    rs = np.random.RandomState(seed)
    h = np.exp(rs.normal(3.0, 0.5, size=n))
    if dyadic_bits is not None:
        h = np.round(h * 2.0 ** dyadic_bits) / 2.0 ** dyadic_bits
        h = np.maximum(h, 2.0 ** -dyadic_bits)
    g = np.maximum(rs.poisson(_levels(n, num_levels) * h), 0).astype(np.float64)
    p = rs.permutation(n)
    return g[p].astype(dtype), h[p].astype(dtype)


def gaussian_mixture(n, num_levels, sigma=5.0, seed=SEED, dtype=np.float32, dyadic_bits=None):
    """b~N(100,5), a~N(q*b,sigma): continuous, tie-free; sums are NOT exact unless dyadic_bits is set."""
    rs = np.random.RandomState(seed)
    b = np.abs(rs.normal(100.0, 5.0, size=n)) + 1.0
    a = rs.normal(_levels(n, num_levels) * b, sigma)
    if dyadic_bits is not None:
        s = 2.0 ** dyadic_bits
        a, b = np.round(a * s) / s, np.maximum(np.round(b * s) / s, 1.0 / s)
    p = rs.permutation(n)
    return a[p].astype(dtype), b[p].astype(dtype)


def uniform_ab(n, seed=SEED, dtype=np.float32, a_lo=1e-5, a_hi=10.0, b_lo=1e-5, b_hi=10.0):
    """The reference timer's inputs (solver_timer.cpp:33-45): a,b ~ U(lo,hi), Poisson keeps a>0."""
    rs = np.random.RandomState(seed)
    a = rs.uniform(a_lo, a_hi, size=n)
    b = rs.uniform(b_lo, b_hi, size=n)
    return a.astype(dtype), b.astype(dtype)


# name -> (n, T, dist, risk_part, generator)   BASELINE.json configs, in order
CONFIGS = {
    "C1": dict(n=1000, T=5, dist=POISSON, risk_part=False,
               gen=lambda dtype, n=1000: poisson_planted(n, 5, seed=SEED + 1, dtype=dtype)),
    "C2": dict(n=2089, T=4, dist=GAUSSIAN, risk_part=True,
               gen=lambda dtype, n=2089: tract_like(n, 4, seed=SEED + 2, dtype=dtype)),
    "C3": dict(n=100000, T=16, dist=POISSON, risk_part=False,
               gen=lambda dtype, n=100000: poisson_planted(n, 16, seed=SEED + 3, dtype=dtype)),
    "C3R": dict(n=100000, T=16, dist=RATIONAL, risk_part=False,
                gen=lambda dtype, n=100000: gaussian_mixture(n, 16, seed=SEED + 3, dtype=dtype)),
    "C4": dict(n=5000, T=8, dist=POISSON, risk_part=False, replicas=4096,
               gen=lambda dtype, n=5000, replicas=4096: poisson_null(n, seed=SEED + 4, dtype=dtype,
                                                                      replicas=replicas)),
    "C5": dict(n=1000000, T=32, dist=POISSON, risk_part=False,
               gen=lambda dtype, n=1000000: poisson_planted(n, 32, seed=SEED + 5, dtype=dtype)),
}


def nominal_cells(n, T):
    """BASELINE.json's unit: n^2 * T."""
    return float(n) * float(n) * float(T)


def actual_cells(n, T):
    """Exact number of (i,k) cells the DP evaluates in levels 2..T (SURVEY.md 8a work counts)."""
    T = min(T, n)
    tot = 0
    for j in range(2, T):
        m = n - j + 1
        tot += m * (m + 1) // 2
    if T >= 2:
        tot += n - T + 1
    return tot
