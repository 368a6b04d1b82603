"""GPU tests of the randomization-test driver (partitionsolvers_b200/randomization.py): the null
score table built from two batched sweep solves on device-generated replicas must equal, bit for bit, what the
reference's procedure computes replica by replica and T by T (SimulatedExperiments.py:74-89)."""
import numpy as np
import pytest

from partitionsolvers_b200 import host, randomization as rz

pytestmark = pytest.mark.gpu


def test_null_table_equals_per_replica_per_T_solves(ctx):
    import torch
    dev = torch.device("cuda", ctx.device)
    cluster_list = (2, 3, 10)
    a_d, b_d = rz.null_replicas_device(7, 300, 100.0, seed=5, device=dev, dtype=np.float32)
    table = rz.solve_scores_device(ctx, a_d, b_d, cluster_list, chunk=4)
    a, b = a_d.cpu().numpy(), b_d.cpu().numpy()
    assert a.min() >= 1 and abs(b.mean() - 100) < 2
    for r in range(7):
        for T in cluster_list:
            rp = host.DPSolver(300, T, a[r], b[r], 1, True, True, dtype=np.float32, ctx=ctx)
            mc = host.DPSolver(300, T, a[r], b[r], 1, False, True, dtype=np.float32, ctx=ctx)
            assert table[f"rp{T}"][r] == rp.get_optimal_score_extern()
            assert table[f"mcd{T}"][r] == mc.get_optimal_score_extern()


def test_thresholds_and_detection_power(ctx):
    cluster_list = (2, 3)
    null = rz.create_null_scores(200, 400, 100.0, cluster_list, ctx=ctx, seed=11)
    assert set(null) == {"rp2", "rp3", "mcd2", "mcd3"} and all(len(v) == 200 for v in null.values())
    assert np.all(null["rp3"] >= null["rp2"])          # a finer partition never scores lower (gtest_all.cpp:561-601)
    th = rz.create_thresholds(null, cluster_list, num_threshold_calculations=50, deviate_size=400, seed=3)
    assert all(np.all(np.diff(v) >= 0) and len(v) == 50 for v in th.values())
    # the null model itself is detected about 5 % of the time, a strongly planted one always
    again = rz.create_null_scores(200, 400, 100.0, cluster_list, ctx=ctx, seed=12)
    p0 = rz.detection_power(again["rp2"], th["rp2"])
    planted = rz.planted_scores(100, 400, 100.0, 2, 0.5, cluster_list, ctx=ctx, seed=13)
    p1 = rz.detection_power(planted["rp2"], th["rp2"])
    assert p0 < 0.2 and p1 == 1.0
