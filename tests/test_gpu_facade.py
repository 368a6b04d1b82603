"""GPU tests of the reference-facing surfaces: the C++ facade (the reference's gtest cases, re-expressed in
tests/cpp/test_facade.cpp) and the pybind11 `proto` module called exactly as solverSWIG_DP.OptimizerSWIG
calls the SWIG one (src/python/solverSWIG_DP.py:57-97)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from partitionsolvers_b200 import host, synth

pytestmark = pytest.mark.gpu

LIBDIR = os.path.join(ROOT, "partitionsolvers_b200", "_lib")


def test_cpp_facade_reference_gtest_cases():
    exe = os.path.join(LIBDIR, "test_facade")
    assert os.path.exists(exe), "run __graft_entry__.build() first"
    env = dict(os.environ, LD_LIBRARY_PATH=LIBDIR + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all checks passed" in r.stdout


@pytest.fixture(scope="module")
def proto():
    sys.path.insert(0, LIBDIR)
    import proto as p
    return p


def test_proto_optimize_one_matches_python_mirror(proto, ctx):
    a, b = synth.poisson_planted(1000, 5, seed=31, dtype=np.float32)
    for dist, rp in ((1, False), (1, True), (0, True), (2, False)):
        subsets, score = proto.optimize_one__DP(1000, 5, a.tolist(), b.tolist(), dist, rp, False, 0.0, 1)
        s = host.DPSolver(1000, 5, a, b, dist, rp, dtype=np.float32, ctx=ctx)
        assert isinstance(subsets, tuple) and isinstance(subsets[0], tuple)
        assert [list(x) for x in subsets] == s.get_optimal_subsets_extern()
        assert np.float32(score) == s.get_optimal_score_extern()


def test_proto_sweeps_and_ols(proto, ctx):
    rs = np.random.RandomState(0xC0FFEE3 % (2 ** 31))
    n = 2000
    b = rs.uniform(90, 110, n)
    a = rs.normal(np.repeat([0.25, 0.75, 1.25, 1.75], n // 4) * b, 1.0)
    allp, t = proto.compute_optimal_num_clusters_OLS_all(n, 12, a.tolist(), b.tolist(), 2, True, False, 0.0, 1)
    assert t == 4 and len(allp) == 13 and all(len(allp[S][0]) == S for S in range(13))
    assert proto.compute_optimal_num_clusters_OLS(n, 12, a.tolist(), b.tolist(), 2, True, False) == 4
    sw = proto.optimize_all__DP(n, 6, a.tolist(), b.tolist(), 2, True, False, 0.0, 1)
    par = proto.sweep_parallel__DP(n, 6, a.tolist(), b.tolist(), 2, True, False, 0.0, 1)
    for S in range(1, 7):
        assert par[S] == sw[S]
    best = proto.sweep_best__DP(n, 6, a.tolist(), b.tolist(), 2, True, False, 0.0, 1)
    assert best[1] == max(sw[S][1] for S in range(2, 7))
    ols = proto.sweep_best_OLS__DP(n, 12, a.tolist(), b.tolist(), 2, True, False, 0.0, 1)
    assert len(ols[0]) == 4
    assert proto.find_optimal_partition__DP(n, 3, a.tolist(), b.tolist(), 2, True, False) == sw[3][0]
    assert proto.find_optimal_score__DP(n, 3, a.tolist(), b.tolist(), 2, True, False) == sw[3][1]


def test_one_pass_sweeps_equal_separate_solves(proto):
    """sweep_parallel__DP / sweep_best__DP take their T results from ONE pass (DPSolver::sweep_extern_scores); the reference
    constructs a solver per T (python_dpsolver.cpp:230-316).  Both objectives, count data with tied priorities, T > n."""
    a, b = synth.poisson_planted(1500, 5, seed=32, dtype=np.float32)
    for dist, rp in ((1, False), (1, True), (0, False), (2, False)):
        par = proto.sweep_parallel__DP(1500, 7, a.tolist(), b.tolist(), dist, rp, False, 0.0, 1)
        assert len(par) == 8 and par[0] == ((), 0.0)
        scores = []
        for S in range(1, 8):
            one = proto.optimize_one__DP(1500, S, a.tolist(), b.tolist(), dist, rp, False, 0.0, 1)
            assert par[S] == one, (dist, rp, S)
            scores.append(one[1])
        best = proto.sweep_best__DP(1500, 7, a.tolist(), b.tolist(), dist, rp, False, 0.0, 1)
        want = max(range(2, 8), key=lambda S: (scores[S - 1], S))  # T..2, strictly better wins: the largest S among equals
        assert best == (par[want][0], scores[want - 1])
    # T capped at n (DP_impl.hpp:37-38): slots above n stay empty, like the reference's results[size()] indexing
    a5, b5 = a[:5], b[:5]
    par = proto.sweep_parallel__DP(5, 8, a5.tolist(), b5.tolist(), 1, True, False, 0.0, 1)
    assert len(par) == 9 and all(len(par[S][0]) == S for S in range(1, 6)) and all(par[S] == ((), 0.0) for S in range(6, 9))
    assert proto.sweep_best__DP(5, 1, a5.tolist(), b5.tolist(), 1, True, False, 0.0, 1)[0] == ()


def test_n_smaller_than_the_vectors_keeps_the_lowest_priorities(proto, oracle, ctx):
    """DPSolver(n, T, a, b) with n < a.size(): the reference sorts the whole vectors and partitions the n elements of lowest
    priority (DP_impl.hpp:10-27).  Tie-free real-valued data: subsets and score equal the reference's exactly."""
    if not oracle.have("ref"):
        pytest.skip("oracle/_ref not built")
    a, b = synth.uniform_ab(500, seed=9, dtype=np.float32)
    for dist, rp in ((2, False), (0, True), (1, True)):
        want_subsets, want_score = oracle.ref_solve_truncated(300, a, b, 4, dist, rp)
        subsets, score = proto.optimize_one__DP(300, 4, a.tolist(), b.tolist(), dist, rp, False, 0.0, 1)
        assert [list(x) for x in subsets] == want_subsets
        assert np.float32(score) == want_score
        s = host.DPSolver(300, 4, a, b, dist, rp, dtype=np.float32, ctx=ctx)
        assert s.get_optimal_subsets_extern() == want_subsets and s.get_optimal_score_extern() == want_score


def test_proto_compute_score_and_errors(proto, oracle):
    a, b = synth.uniform_ab(300, seed=2, dtype=np.float32)
    got = proto.compute_score(a.tolist(), b.tolist(), 2, False, False)
    assert np.float32(got) == oracle.range_score(a, b, 2, False, 0, 300, dtype=np.float32)
    with pytest.raises(proto.distributionException):
        proto.optimize_one__DP(300, 3, a.tolist(), b.tolist(), 5, True, False)


def test_ltss_matches_oracle_and_dp(oracle, ctx, proto):
    """LTSSSolver<float> on the device == the oracle restatement of LTSS_impl.hpp:73-111 (bit for bit), and the
    reference's cross-solver property TestHighestScoringSetOf2TieOutAllDists (gtest_all.cpp:500-559): the top
    subset of the T=2 multiple-clustering partition is the LTSS subset."""
    rs = np.random.RandomState(33)
    for dist in (0, 1, 2):
        for n in (1, 2, 37, 300, 4097):
            a = (rs.uniform(0.1, 10, n) if dist == 1 else rs.uniform(-10, 10, n)).astype(np.float32)
            b = rs.uniform(0.1, 10, n).astype(np.float32)
            want = oracle.ltss("port", a, b, dist)
            got = ctx.ltss(a, b, dist)
            assert np.array_equal(got["perm"], want["perm"])
            assert got["subset"].tolist() == want["subset"].tolist(), (dist, n)
            assert got["score"] == want["score"]
            sub, sc = proto.optimize_one__LTSS(n, a.tolist(), b.tolist(), dist)
            assert list(sub) == want["subset"].tolist() and np.float32(sc) == want["score"]
            assert list(proto.find_optimal_partition__LTSS(n, a.tolist(), b.tolist(), dist)) == list(sub)
            assert proto.find_optimal_score__LTSS(n, a.tolist(), b.tolist(), dist) == sc
    for dist in (0, 1):
        a = rs.uniform(0.1, 10, 500).astype(np.float32)
        b = rs.uniform(0.1, 10, 500).astype(np.float32)
        dp = host.DPSolver(500, 2, a, b, dist, False, True, dtype=np.float32, ctx=ctx)
        assert sorted(dp.subsets[1].tolist()) == sorted(ctx.ltss(a, b, dist)["subset"].tolist())


@pytest.mark.parametrize("prog,args", [("ref_DP_solver_demo", []), ("ref_LTSS_solver_demo", []),
                                       ("ref_solver_timer", ["400", "20", "100", "5"])])
def test_reference_programs_run_on_the_engine(prog, args):
    """The reference's own demo / timer sources, compiled unchanged against the facade (build.py:
    build_reference_programs) and executed on the B200.  Skipped where the binaries were not built (no reference
    checkout at build time)."""
    exe = os.path.join(LIBDIR, prog)
    if not os.path.exists(exe):
        pytest.skip(f"{prog} not built")
    r = subprocess.run([exe] + args, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    if prog == "ref_solver_timer":  # CSV of constructor times in microseconds, (n+1) rows x (T+1) columns
        rows = [l for l in r.stdout.strip().splitlines() if "," in l]
        assert len(rows) == 401 and all(len(l.rstrip(",").split(",")) == 21 for l in rows)
        # sizes visited: n = 20, 120, 220, 320 (T + k*NStride) and T = 5, 10, 15, 20 (solver_timer.cpp:60-64)
        assert int(rows[320].rstrip(",").split(",")[20]) > 0 and int(rows[120].rstrip(",").split(",")[5]) > 0
        # ... and the file is what the reference's plotting script consumes (src/python/plot_times.py:83-91): read_csv
        # without a header, the trailing-delimiter column dropped, then the visited (n, T) grid selected by label
        import io
        import pandas as pd
        n_, T_, NStride, TStride = 400, 20, 100, 5
        df = pd.read_csv(io.StringIO("\n".join(rows)), sep=",", header=None, index_col=None)
        df = df.drop(columns=[df.columns[-1]])
        grid = df.loc[range(T_, n_ + 1, NStride), range(TStride, T_ + 1, TStride)]
        assert grid.shape == (4, 4)
        for nn in grid.index:
            for tt in grid.columns:  # the timer only runs sampleSize > numParts (solver_timer.cpp:61)
                assert (grid.loc[nn, tt] > 0) == (nn > tt), (nn, tt)
