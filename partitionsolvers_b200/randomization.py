"""Randomization-test driver over the batched engine.

The reference's experiment script (src/python/SimulatedExperiments.py) builds a table of NULL scores -- for every
Monte-Carlo replica b~Poisson(lam), a~Poisson(b) it solves the risk-partitioning and the multiple-clustering
problem for each T in CLUSTER_LIST (:74-89), spread over a multiprocessing pool (:178-222) -- then bootstraps
95 % thresholds from that table (:47-52, :224-236) and scores planted instances against them (:304-333).

Here the replicas are generated ON THE DEVICE (torch's Philox stream; nothing is copied in), the whole table comes
from TWO batched engine calls (one per objective): a single DP pass to T = max(CLUSTER_LIST) with sweep=1 yields
the optimal S-partition for every S <= T, and the reference's own tests pin "sweep entry S == standalone solve with
T=S" (gtest_all.cpp:603-662).  Per-replica scores are finished on the host exactly as the reference's getters do
(DP_impl.hpp:131-146, 232-259, 267-276), in DataType = float like its Python bindings.
"""
import math

import numpy as np

from . import capi

POISSON = 1


def _objective_rp(sc, S, dtype):
    """optimal_score_ of the S-partition: scores accumulated left to right in DataType (DP_impl.hpp:137)."""
    obj = np.zeros(sc.shape[0], dtype)
    for t in range(S):
        obj = (obj + sc[:, t]).astype(dtype)
    return obj


def _objective_mc(sc, S, dtype):
    """get_optimal_score_extern for multiple clustering: subsets reordered by ascending score (stable), the lowest
    one dropped, the rest accumulated in double (DP_impl.hpp:232-259, 267-276)."""
    srt = np.sort(sc[:, :S], axis=1, kind="stable")
    acc = np.zeros(sc.shape[0], np.float64)
    for t in range(1, S):
        acc = acc + srt[:, t].astype(np.float64)
    return acc.astype(dtype)


def null_replicas_device(num, n, poisson_intensity, seed, device, dtype):
    """b ~ Poisson(lam), a ~ Poisson(b) (SimulatedExperiments.py:79-80), generated in HBM."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    lam = torch.full((num, n), float(poisson_intensity), device=device, dtype=tdt)
    b = torch.poisson(lam, generator=g).clamp_(min=1.0)
    a = torch.poisson(b, generator=g).clamp_(min=1.0)  # Poisson scores need a > 0 (solver_timer.cpp:35-38)
    return a.contiguous(), b.contiguous()


def solve_scores_device(ctx, a_d, b_d, cluster_list, dtype=np.float32, dist=POISSON, chunk=2048):
    """Score columns rp<T>, mcd<T> for device-resident instances a_d, b_d of shape [R, n]."""
    import torch
    R, n = a_d.shape
    Tmax = min(max(cluster_list), n)
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    out = {}
    for rp, prefix in ((True, "rp"), (False, "mcd")):
        cols = {T: np.empty(R, dtype) for T in cluster_list}
        for lo in range(0, R, chunk):
            hi = min(R, lo + chunk)
            Rc = hi - lo
            sc_d = torch.empty((Rc, Tmax + 1, Tmax), dtype=tdt, device=a_d.device)
            ctx.wait_torch(a_d.device)  # the engine's own stream is not ordered with torch's: the replicas must have been drawn
            ctx.solve_device(a_d[lo:hi].data_ptr(), b_d[lo:hi].data_ptr(), n, Tmax, dist, rp, dtype, R=Rc, sweep=True,
                             d_scores=sc_d.data_ptr())
            sc = sc_d.cpu().numpy()
            for T in cluster_list:
                S = min(T, n)
                cols[T][lo:hi] = _objective_rp(sc[:, S, :], S, dtype) if rp else _objective_mc(sc[:, S, :], S, dtype)
        for T in cluster_list:
            out[f"{prefix}{T}"] = cols[T]
    return out


def create_null_scores(num_null_experiments, deviate_size, poisson_intensity, cluster_list, ctx=None, seed=1869,
                       dtype=np.float32):
    """The reference's Baselines.create_null_scores (:203-222) as two batched solves.  Returns a dict of columns."""
    import torch
    ctx = ctx or capi.Context()
    dev = torch.device("cuda", ctx.device)
    a_d, b_d = null_replicas_device(num_null_experiments, deviate_size, poisson_intensity, seed, dev, dtype)
    return solve_scores_device(ctx, a_d, b_d, cluster_list, dtype=dtype)


def bootstrap_95th_percentile(null_scores, num_threshold_calculations, deviate_size, rng):
    """SimulatedExperiments.py:47-52."""
    null_scores = np.asarray(null_scores)
    thresholds = np.zeros(num_threshold_calculations)
    for i in range(num_threshold_calculations):
        thresholds[i] = np.quantile(rng.choice(null_scores, size=deviate_size, replace=True), 0.95)
    return np.sort(thresholds)


def create_thresholds(null_scores, cluster_list, num_threshold_calculations, deviate_size, seed=1869):
    """Baselines.create_thresholds (:224-236): one sorted threshold vector per score column."""
    rng = np.random.RandomState(seed)
    return {col: bootstrap_95th_percentile(null_scores[col], num_threshold_calculations, deviate_size, rng)
            for T in cluster_list for col in (f"rp{T}", f"mcd{T}")}


def planted_levels(n, num_true_clusters, epsilon):
    """q of QATask (:128-146, the non-figure branch): equal blocks with rates linspace(eps/k, 2-eps/k, k)."""
    split = n // num_true_clusters
    resid = n - split * num_true_clusters
    splits = [split + (1 if i < resid else 0) for i in range(num_true_clusters)]
    levels = np.linspace(epsilon / num_true_clusters, 2 - epsilon / num_true_clusters, num_true_clusters)
    return np.concatenate([np.full(s, l) for s, l in zip(splits, levels)])


def detection_power(scores, thresholds):
    """QA.compute_detection_power (:304-333) for one method and one epsilon: the fraction of bootstrap thresholds
    each score exceeds, averaged over the scores."""
    thresholds = np.asarray(thresholds)
    lo, hi = thresholds.min(), thresholds.max()
    power = 0.0
    for k in np.asarray(scores, dtype=np.float64):
        if k <= lo:
            continue
        if k >= hi:
            power += 1.0
        else:
            power += float(np.sum(k > thresholds)) / len(thresholds)
    return power / max(1, len(scores))


def planted_scores(num_experiments, n, poisson_intensity, num_true_clusters, epsilon, cluster_list, ctx=None,
                   seed=1, dtype=np.float32):
    """Scores of planted instances (QATask's data model) through the same batched path; rp columns only are used by
    the reference's power computation, both are returned."""
    import torch
    ctx = ctx or capi.Context()
    dev = torch.device("cuda", ctx.device)
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    lam = torch.full((num_experiments, n), float(poisson_intensity), device=dev, dtype=tdt)
    b = torch.poisson(lam, generator=g).clamp_(min=1.0)
    q = torch.as_tensor(planted_levels(n, num_true_clusters, epsilon), device=dev, dtype=tdt)
    a = torch.poisson(b * q, generator=g).clamp_(min=1.0)
    return solve_scores_device(ctx, a.contiguous(), b.contiguous(), cluster_list, dtype=dtype)
