"""GPU test of the row-sharded building blocks (ps_shard_*): a world of 2 and 3 ranks emulated on ONE device
(the ranks' kernels run one after another, the all-gather is a concatenation), compared bit-for-bit with the
unsharded solve.  The NCCL all-gather itself is exercised by bench.py --workload sharded on a multi-GPU box."""
import numpy as np
import pytest

from partitionsolvers_b200 import capi, multigpu, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [3000, 7000])  # 7000: the shards run the screened level kernel (csrc/screen.cuh)
@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_sharded_levels_equal_unsharded(ctx, world, dt, n):
    import torch
    dev = torch.device("cuda", ctx.device)
    T, dist_id, rp = 6, 1, False
    a, b = synth.poisson_planted(n, 6, seed=41, dtype=dt)
    full = ctx.solve_raw(a, b, T, dist_id, rp, dtype=dt, levels=True)
    engines = [multigpu.CudaShardEngine(ctx, dev) for _ in range(world)]
    perm, a_s, b_s = engines[0].sort(a, b, dt)
    assert np.array_equal(perm, full["perm"])
    for r, e in enumerate(engines):
        e.create(a_s, b_s, n, T, dist_id, rp, dt, r, world)
    assert all(e.screened == (n >= 6000) for e in engines)
    chunks = [e.new_chunk() for e in engines]

    def exchange():
        g = torch.cat(chunks)
        return g, multigpu.natural_from_gathered(g.cpu().numpy(), n, world, engines[0].chunk_rows)

    for e, c in zip(engines, chunks):
        e.level1(c)
    g, nat = exchange()
    assert np.array_equal(nat, full["max_score"][1])
    for j in range(2, T + 1):
        for e, c in zip(engines, chunks):
            e.set_prev(g)
            e.level(j, c)
        g, nat = exchange()
        nrows = 1 if j == T else n - j + 1
        assert np.array_equal(nat[:nrows], full["max_score"][j][:nrows]), j
        for i in (0, 1, 127, 128, 129, 255, 256, nrows - 1):
            if i < nrows:
                own = multigpu.owner_of_row(i, world)
                assert engines[own].next_start(j, i) == full["next_start"][j][i]
                if own != 0:
                    assert engines[0].next_start(j, i) == -2
    for e in engines:
        e.close()


@pytest.mark.parametrize("kind", ["counts", "real"])
def test_sharded_solver_subset_scores(ctx, kind):
    """ShardedSolver end to end on one rank: boundaries, subsets and the per-subset scores (ps_shard_subset_scores: one
    thread per subset; prefix differences when the shard runs on exactly summable inputs, a left-to-right walk otherwise)
    equal the unsharded solve bit for bit."""
    import torch
    dev = torch.device("cuda", ctx.device)
    n, T = 7000, 6
    if kind == "counts":
        a, b = synth.poisson_planted(n, 6, seed=43, dtype=np.float64)
    else:
        a, b = synth.tract_like(n, 4, seed=44, dtype=np.float64)
        a = np.maximum(a, 1.0)
    full = ctx.solve_raw(a, b, T, 1, False, dtype=np.float64)
    sh = multigpu.ShardedSolver(multigpu.CudaShardEngine(ctx, dev), a, b, T, 1, False, dtype=np.float64)
    assert np.array_equal(sh.bounds, full["bounds"].ravel())
    assert np.array_equal(sh.perm, full["perm"])
    assert np.array_equal(sh.subset_scores, full["subset_scores"].ravel())


# ------------------------------------------------------------------------------------------------------------------
# ps_solve_sharded_*: the per-level loop, the NCCL all-gather and the scatter inside the library (include/ps_b200.h)
# ------------------------------------------------------------------------------------------------------------------
SHARD_CASES = [
    ("counts_screened", 8192, 6, 1, False),   # exact-sum screened levels (live tiles, fold -> packed chunk)
    ("counts_small", 1000, 5, 1, False),      # below the screening threshold: exhaustive kernel, prefix-free running sums
    ("real_running", 6500, 5, 2, False),      # real-valued data: RUNNING sums, sharded level-1 walk; exhaustive kernel, and
                                              # (audited, screen = 2) the running-sum screens of screen_run.cuh
    ("counts_rp", 6500, 4, 1, True),          # Poisson risk-part: box tiers
]


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("name,n,T,dist_id,rp", SHARD_CASES)
def test_solve_sharded_one_rank_equals_unsharded(ctx, dt, name, n, T, dist_id, rp):
    """A communicator of ONE rank on this device drives the complete sharded path -- owned-row kernels, ncclAllGather on the
    engine's stream, scatter, replicated last level and backtrace -- and must return the bytes of ps_solve_*: sort index,
    boundaries, subset scores, every maxScore_ / nextStart_ cell, with and without the sweep."""
    if name.startswith("counts"):
        a, b = synth.poisson_planted(n, T, seed=51, dtype=dt)
    else:
        a, b = synth.gaussian_mixture(n, T, seed=52, dtype=dt)
    comm = ctx.comm_create(ctx.comm_unique_id(), 0, 1)
    try:
        for sweep, scr in ((False, 0), (True, 0), (False, 2)):  # screen = 2: audited screens (running-sum ones at any size)
            full = ctx.solve_raw(a, b, T, dist_id, rp, dtype=dt, levels=True, sweep=sweep, screen=scr)
            used = ctx.timings()["screen_used"]
            sh = ctx.solve_sharded(comm, a, b, T, dist_id, rp, dtype=dt, levels=True, sweep=sweep, screen=scr)
            t = ctx.timings()
            assert t["screen_used"] == used and t["comm_bytes"] > 0 and t["screen_violations"] == 0
            if scr == 2 and name == "real_running":
                assert used == 2, "the running-sum screen did not run in the sharded solve"
            for k in ("perm", "bounds", "subset_scores", "max_score", "next_start"):
                assert np.array_equal(np.asarray(sh[k]).ravel(), np.asarray(full[k]).ravel()), (name, k, sweep, scr)
    finally:
        ctx.comm_destroy(comm)


def test_solve_sharded_two_processes():
    """Two ranks, two GPUs, NCCL between them (tools/shard_check.py under torchrun): every rank's result equals the
    single-GPU solve.  Skipped on a one-GPU box; the 8-GPU run is bench.py --gpus 8 (`sharded`)."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(root, "tools", "shard_check.py"),
                        "--size", "20000", "--parts", "6", "--check"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "SHARD_CHECK_OK" in r.stdout
    # real-valued data: running sums (level-1 walk and checkpoints for owned rows only, row 0's level T on rank 0), the
    # exhaustive kernel and -- audited -- the running-sum screen, both precisions
    for extra in (["--data", "real", "--dtype", "f64"], ["--data", "real", "--dtype", "f32", "--screen", "2"],
                  ["--data", "real", "--dtype", "f64", "--screen", "2"]):
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                            "--master-addr", "127.0.0.1", "--master-port", "29518", os.path.join(root, "tools", "shard_check.py"),
                            "--size", "12000", "--parts", "5", "--check"] + extra, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "SHARD_CHECK_OK" in r.stdout, (extra, r.stdout[-3000:] + r.stderr[-3000:])
