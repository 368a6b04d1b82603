"""pytest configuration: the `gpu` marker, checker builds, golden-fixture loading, comparison helpers."""
import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU checkers (oracle/): the port is (re)built on demand; the reference shim where possible."""
    from oracle import oracle_py as O
    O.build(("port",))
    if os.path.isdir("/root/reference/src/cpp/include"):
        O.build(("ref",))
    return O


@pytest.fixture(scope="session")
def ctx():
    """One CUDA context for the GPU tests.  No fallback: missing library or device is a failure."""
    from partitionsolvers_b200 import capi
    c = capi.Context()
    yield c
    c.close()


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    g["dtype"] = g["a"].dtype.type
    subsets, off = [], 0
    for s in g["subset_sizes"]:
        subsets.append(g["subset_elems"][off:off + s])
        off += s
    g["subsets"] = subsets
    if int(g["sweep"]):
        allp = []
        Tc = g["all_sizes"].shape[1]
        for S in range(Tc + 1):
            off, ss = 0, []
            for q in range(S):
                ss.append(g["all_elems"][S, off:off + g["all_sizes"][S, q]])
                off += g["all_sizes"][S, q]
            allp.append((ss, g["all_scores"][S]))
        g["all"] = allp
    return g


def assert_same_partition(got, want):
    assert len(got) == len(want)
    for x, y in zip(got, want):
        assert np.array_equal(np.asarray(x), np.asarray(y))


_LOG_EXACT = {}


def device_log_matches_host(dtype):
    """True when the device restatement of glibc's log/logf (csrc/glibc_log.cuh) reproduces THIS host's libm
    bit for bit on a probe set -- the case on every x86-64 host whose glibc selects the FMA variants, i.e. all
    AVX2-class CPUs.  Then the Poisson scores are bit-identical to the oracle's and no tolerance applies."""
    key = np.dtype(dtype).name
    if key not in _LOG_EXACT:
        from oracle import oracle_py as O
        from partitionsolvers_b200 import capi
        rs = np.random.RandomState(5)
        x = np.concatenate([1.0 + rs.uniform(-0.07, 0.07, 20000), rs.uniform(0.01, 50.0, 20000)]).astype(dtype)
        c = capi.Context()
        got, _ = c.eval_log(x, dtype=dtype)
        c.close()
        _LOG_EXACT[key] = bool(np.array_equal(got, O.host_log(x, dtype=dtype)))
    return _LOG_EXACT[key]


def tol_for(dtype, uses_log):
    """Parity tolerances (DESIGN.md "Parity").  RUNNING sums: Gaussian and Rational scores are bit-exact; so
    are the Poisson scores wherever the host's libm is the one csrc/glibc_log.cuh restates (checked at run
    time).  Only on a host with a different libm does the Poisson path fall back to "<= 1 ulp of log"
    tolerances (amplified by the cancellation in A*log(A/B)+B-A)."""
    if not uses_log or device_log_matches_host(dtype):
        return 0.0
    return 1e-9 if dtype == np.float64 else 2e-4


def check_levels(ms_g, ns_g, ms_o, ns_o, rtol, tie_rtol, recompute=None):
    """maxScore_ rows within rtol (relative to the cell, floor = rtol * largest finite value of its
    level row); nextStart_ identical except where the two candidate splits tie within tie_rtol under
    the ORACLE's arithmetic (north-star: bit-exact splits except where the objective ties).
    Returns the number of flipped-but-tied splits."""
    assert ms_g.shape == ms_o.shape and ns_g.shape == ns_o.shape
    if rtol == 0.0:
        assert np.array_equal(ms_g, ms_o), f"{int((ms_g != ms_o).sum())} maxScore_ cells differ (bit-exact expected)"
        assert np.array_equal(ns_g, ns_o), f"{int((ns_g != ns_o).sum())} nextStart_ cells differ"
        return 0
    mo = ms_o.astype(np.float64)
    mg = ms_g.astype(np.float64)
    low = float(np.finfo(ms_o.dtype).min)
    written = mo > low / 2
    assert np.array_equal(written, mg > low / 2), "different sets of written cells"
    rowmax = np.max(np.where(written, np.abs(mo), 0.0), axis=1, keepdims=True)
    err = np.where(written, np.abs(mg - mo), 0.0)
    lim = rtol * np.maximum(np.abs(mo), rowmax)
    bad = written & (err > lim)
    assert not bad.any(), (f"maxScore_: {int(bad.sum())} cells beyond rtol {rtol}; worst "
                           f"{float(np.max(err / np.maximum(lim / rtol, 1e-300))):.3e}")
    diff = np.argwhere(ns_g != ns_o)
    for j, i in diff:
        assert abs(mg[j, i] - mo[j, i]) <= tie_rtol * max(abs(mo[j, i]), float(rowmax[j, 0])), \
            (int(j), int(i), int(ns_g[j, i]), int(ns_o[j, i]))
        if recompute is not None:
            vg, vo = recompute(int(j), int(i), int(ns_g[j, i])), recompute(int(j), int(i), int(ns_o[j, i]))
            assert abs(vg - vo) <= tie_rtol * max(abs(vo), float(rowmax[j, 0])), (int(j), int(i), vg, vo)
    return len(diff)
