"""Python mirror of the host-side facade (include/partitionsolvers/DP.hpp).

The CUDA engine returns, per instance, the sort index, the subset boundaries in sorted order and the
per-subset range scores.  Everything the reference does AFTER the DP fill is O(n + T) host work and
is finished here with the reference's own semantics:

  optimize_for_fixed_S  DP_impl.hpp:122-153   subsets = perm[b_t:b_{t+1}], objective accumulated in
                                              DataType, regulariser gamma*pow(S,reg_power) in double
  reorder_subsets       DP_impl.hpp:232-259   mult-clust only: stable sort by ascending subset score
  optimize              DP_impl.hpp:200-230   sweep table of T+1 entries (entry 0 empty)
  find_optimal_t        DP_impl.hpp:155-198   OLS elbow on log score differences (float, Eigen-free)
  getters               DP_impl.hpp:261-294   incl. the mult-clust quirk of get_optimal_score_extern
"""
import math

import numpy as np

from . import capi

_default_ctx = None


def default_context():
    """Lazily created process-wide context (CUDA is never touched at import time)."""
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = capi.Context()
    return _default_ctx


def ols_num_clusters(scores, dtype):
    """DP_impl.hpp:155-198 without Eigen: regress log(s[S]-s[S-1]) on log(S+1), S=2..T, in float."""
    m = len(scores) - 2
    if m <= 0:
        return 1
    sc = np.asarray(scores, dtype=dtype)
    with np.errstate(all="ignore"):
        x = np.log(np.arange(3, m + 3).astype(dtype)).astype(np.float32)
        y = np.log((sc[2:] - sc[1:-1]).astype(dtype)).astype(np.float32)
    xd, yd = x.astype(np.float64), y.astype(np.float64)
    sx, sy, sxx, sxy = xd.sum(), yd.sum(), (xd * xd).sum(), (xd * yd).sum()
    with np.errstate(all="ignore"):
        den = m * sxx - sx * sx
        if den != 0:
            b0 = np.float32((m * sxy - sx * sy) / den)
            b1 = np.float32((sy - np.float64(b0) * sx) / m)
        else:  # one sample (T = 2): basic solution of the under-determined system, residual 0 -> the choice is 1
            b0 = np.float32(yd[0] / xd[0])
            b1 = np.float32(0.0)
        best, minres = 0, np.finfo(dtype).max
        for i in range(m):
            r = np.float32(np.float32(y[i] - np.float32(b0 * x[i])) - b1)
            if r < minres:
                best, minres = i, r
    return best + 1


def _partition(perm, bounds_row, scores_row, S, risk_part, gamma, reg_power, dtype):
    subsets, sc = [], []
    obj = dtype(0)
    for t in range(S):
        lo, hi = int(bounds_row[t]), int(bounds_row[t + 1])
        subsets.append(perm[lo:hi].copy() if 0 <= lo <= hi else np.empty(0, np.int32))
        s = dtype(scores_row[t])
        sc.append(s)
        obj = dtype(obj + s)
    if not risk_part:
        order = sorted(range(S), key=lambda q: _lt_key(sc[q]))  # Python's sort is stable
        subsets = [subsets[q] for q in order]
        sc = [sc[q] for q in order]
    obj = dtype(float(obj) - float(dtype(gamma)) * math.pow(S, reg_power))
    return subsets, np.array(sc, dtype=dtype), obj


def _lt_key(v):
    # comparator of reorder_subsets is '<' on DataType; NaN compares false both ways
    return float(v) if v == v else float("inf")


class DPSolver:
    """Same observable behaviour as the reference's DPSolver<DataType> (DP.hpp:35-198)."""

    def __init__(self, n, T, a, b, parametric_dist=0, risk_partitioning_objective=False,
                 use_rational_optimization=True, gamma=0.0, reg_power=1, sweep_down=False,
                 find_optimal_t=False, dtype=np.float32, ctx=None, sum_mode=capi.PS_SUM_RUNNING,
                 perm=None, levels=False, chunk_len=0):
        del use_rational_optimization  # accepted and ignored, like the reference
        a = np.asarray(a, dtype=dtype)
        b = np.asarray(b, dtype=dtype)
        n = min(int(n), a.shape[0])  # DP_impl.hpp:35
        self.dtype = dtype
        self.ctx = ctx or default_context()
        self.risk_part = bool(risk_partitioning_objective)
        self.gamma, self.reg_power = gamma, int(reg_power)
        sweep = bool(sweep_down or find_optimal_t)
        full_index = None
        if n < a.shape[0]:
            # the reference sorts the WHOLE vectors and keeps the first n sorted elements (DP_impl.hpp:10-27): stable order on
            # the host, kept elements handed to the engine already sorted
            with np.errstate(all="ignore"):
                key = a / b
            full_index = np.argsort(np.where(key == 0, dtype(0), key), kind="stable").astype(np.int32)
            a, b = a[full_index[:n]], b[full_index[:n]]
            perm = np.arange(n, dtype=np.int32)
        raw = self.ctx.solve_raw(a[:n], b[:n], T, int(parametric_dist), self.risk_part, dtype=dtype,
                                 sum_mode=sum_mode, perm=perm, sweep=sweep, levels=levels,
                                 chunk_len=chunk_len)
        if full_index is not None:
            raw["perm"] = full_index
        self.raw = raw
        self.n, self.T = n, raw["T"]
        Tc = self.T
        perm_out = raw["perm"]
        self.priority_sortind = perm_out
        self.optimal_num_clusters_OLS = 0
        self.all = None
        if sweep:
            self.all = [([], dtype(0))]
            sbs_T = None
            for S in range(1, Tc + 1):
                subsets, sc, obj = _partition(perm_out, raw["bounds"][S], raw["subset_scores"][S], S,
                                              self.risk_part, gamma, self.reg_power, dtype)
                self.all.append((subsets, obj))
                if S == Tc:
                    sbs_T = sc
            self.score_by_subset = sbs_T
            self.subsets = [np.empty(0, np.int32) for _ in range(Tc)]  # create() defaults, DP_impl.hpp:86
            self.optimal_score = dtype(0)
            if find_optimal_t:
                t = ols_num_clusters([s for _, s in self.all], dtype)
                self.optimal_num_clusters_OLS = t
                self.subsets, self.optimal_score = self.all[t]
        else:
            self.subsets, self.score_by_subset, self.optimal_score = _partition(
                perm_out, raw["bounds"][0], raw["subset_scores"][0], Tc, self.risk_part, gamma,
                self.reg_power, dtype)

    # -- getters, DP_impl.hpp:261-294
    def get_optimal_subsets_extern(self):
        return [s.tolist() for s in self.subsets]

    def get_optimal_score_extern(self):
        if self.risk_part:
            return self.dtype(self.optimal_score)
        acc = 0.0
        for v in self.score_by_subset[1:]:
            acc += float(v)
        return self.dtype(acc - float(self.dtype(self.gamma)) * math.pow(self.T, self.reg_power))

    def get_score_by_subset_extern(self):
        return self.score_by_subset.copy()

    def get_all_subsets_and_scores_extern(self):
        return [([s.tolist() for s in ss], sc) for ss, sc in (self.all or [])]

    def get_optimal_num_clusters_OLS_extern(self):
        return self.optimal_num_clusters_OLS
