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
